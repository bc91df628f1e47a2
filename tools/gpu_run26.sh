#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" > $OUT/summary.txt; tail -2 $OUT/smoke.log >> $OUT/summary.txt
timeout -k 5 900 python -m pytest tests -q -m gpu --timeout=600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/summary.txt
tail -3 $OUT/pytest_gpu.log >> $OUT/summary.txt
for G in 1 0; do
TNB_GRAPHS=$G timeout -k 5 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_g$G.log 2>&1; echo "bench graphs=$G rc=$?" >> $OUT/summary.txt
tail -1 $OUT/bench_g$G.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1), d['cuda_graphs'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
done
timeout -k 5 600 python tools/torch_cuda_baseline.py 10 5 > $OUT/torch_cuda.log 2>&1
grep impl $OUT/torch_cuda.log | cut -c1-150 >> $OUT/summary.txt
cat $OUT/summary.txt
