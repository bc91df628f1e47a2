#!/bin/bash
# usage: tools/gpu_exp.sh <outdir-name>: wgrad fill-mapping experiment (tools/exp_wgrad_map.py) + tests and bench per mapping
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 150 python tools/exp_wgrad_map.py generic > $OUT/generic.log 2>&1; echo "generic rc=$?" > $OUT/summary.txt
for m in 0 4 8; do TNB_WGRAD_MAP=$m timeout -k 5 60 python tools/exp_wgrad_map.py stacked >> $OUT/stacked.log 2>&1; done
for m in 4 8; do
  TNB_WGRAD_MAP=$m timeout -k 5 120 python -m pytest tests/test_gpu_conv.py -x -q -k wgrad > $OUT/pytest_map$m.log 2>&1; echo "pytest map$m rc=$?" >> $OUT/summary.txt
  tail -1 $OUT/pytest_map$m.log >> $OUT/summary.txt
done
for m in 0 4 8; do
  TNB_WGRAD_MAP=$m timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_map$m.log 2>&1
  tail -1 $OUT/bench_map$m.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('map $m: ms',round(d['ms_per_step'],3),'wgrad',round(d['kernel_breakdown']['conv3x3 wgrad']['ms_per_step'],3),'clk',d['clocks']['sm_mhz'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
done
cat $OUT/generic.log $OUT/stacked.log $OUT/summary.txt
