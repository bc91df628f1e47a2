#!/bin/bash
# usage: tools/gpu_r2_tma_input.sh <outdir-name>: tensor-TMA staging of the network input (first convolution):
# op-level parity first (own short timeout: a wrong tensor map shows up as an mbarrier timeout trap), the whole GPU suite,
# then the per-launch A/B/A on one box against the gather path (TNB_INPUT_TMA=0)
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 120 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k tensor_tma > $OUT/pytest_tma.log 2>&1; echo "pytest tensor_tma rc=$?" > $OUT/summary.txt
tail -4 $OUT/pytest_tma.log | cut -c1-300 >> $OUT/summary.txt
if grep -q passed $OUT/pytest_tma.log && ! grep -q failed $OUT/pytest_tma.log; then
  timeout -k 5 900 python -m pytest tests -x -q -m gpu --timeout=600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/summary.txt
  tail -6 $OUT/pytest_gpu.log | cut -c1-300 >> $OUT/summary.txt
  for v in 1 0 1; do
    TNB_INPUT_TMA=$v timeout -k 5 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-alt-precision --per-launch 2> $OUT/launches_tma$v.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('TNB_INPUT_TMA=$v: ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),{k:round(x['ms_per_step'],3) for k,x in d['kernel_breakdown'].items()},'clk',d['clocks']['sm_mhz'])" >> $OUT/summary.txt
    grep "launch   0 " $OUT/launches_tma$v.txt >> $OUT/summary.txt
  done
fi
cat $OUT/summary.txt
