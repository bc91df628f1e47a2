#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
for P in 0 1; do
  TNB_CONV_PLAN=$P timeout 300 python tools/ablate_plan.py >> $OUT/plan.log 2>&1
  for V in 0 64; do
    echo "PLAN=$P variant=$V" >> $OUT/bench.log
    TNB_CONV_PLAN=$P timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variant $V 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3)); 
for k,v in d['kernel_breakdown'].items(): print('   ',k, round(v['ms_per_step'],3))" >> $OUT/bench.log 2>&1
  done
done
timeout 600 python tools/bench_configs.py > $OUT/configs.log 2>&1
cat $OUT/plan.log $OUT/bench.log; cat $OUT/configs.log | cut -c1-400
