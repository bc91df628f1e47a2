#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 900 python -m pytest tests -q -m gpu --timeout=600 -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $OUT/summary.txt
tail -3 $OUT/pytest_gpu.log >> $OUT/summary.txt
for CA in 0 1; do
  echo "CPASYNC_CA=$CA" >> $OUT/summary.txt
  TNB_CPASYNC_CA=$CA timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1));
for k,v in d['kernel_breakdown'].items(): print('   ',k, round(v['ms_per_step'],3))" >> $OUT/summary.txt 2>&1
done
TNB_CPASYNC_CA=0 timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py 1 1 > $OUT/ncu_launches.log 2>&1
TNB_CPASYNC_CA=1 timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_ca.csv python tools/profile_step.py 1 1 > $OUT/ncu_launches_ca.log 2>&1
cat $OUT/summary.txt
