#!/bin/bash
# usage: tools/gpu_exp.sh <outdir-name>: wgrad kernels alone (tools/ablate_wgrad.py), conv parity tests, bench
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 120 python -m pytest tests/test_gpu_conv.py -x -q -k "wgrad" > $OUT/pytest_wgrad.log 2>&1; echo "pytest wgrad rc=$?" > $OUT/summary.txt
tail -15 $OUT/pytest_wgrad.log | cut -c1-300 >> $OUT/summary.txt
timeout -k 5 100 python tools/ablate_wgrad.py > $OUT/ablate_wgrad.log 2>&1; echo "ablate rc=$?" >> $OUT/summary.txt
grep -E "shape|full|no-MMA|no-fill" $OUT/ablate_wgrad.log >> $OUT/summary.txt
timeout -k 5 200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_tracknet.py -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/summary.txt
tail -3 $OUT/pytest.log | cut -c1-300 >> $OUT/summary.txt
for pair in 1 0; do
TNB_WGRAD_PAIR=$pair timeout -k 5 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_pair$pair.log 2>&1
tail -1 $OUT/bench_pair$pair.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('bench pair=$pair: ms',round(d['ms_per_step'],3),'fps',round(d['value'],1),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
done
cat $OUT/summary.txt
