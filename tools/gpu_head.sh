#!/bin/bash
# usage: tools/gpu_head.sh <outdir-name>: short HEAD validation (gpu tests, smoke, bench, configs) for a tight GPU budget
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 420 python -m pytest tests -x -q -m gpu --timeout=300 --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $OUT/summary.txt
tail -12 $OUT/pytest_gpu.log >> $OUT/summary.txt
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary.txt; tail -1 $OUT/smoke.log >> $OUT/summary.txt
timeout -k 5 240 python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/summary.txt
tail -1 $OUT/bench.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1), d['cuda_graphs'], d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
timeout -k 5 200 python tools/bench_configs.py > $OUT/configs.log 2>&1; echo "configs rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt; cut -c1-1200 $OUT/configs.log | head -1
