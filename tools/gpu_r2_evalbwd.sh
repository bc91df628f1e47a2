#!/bin/bash
# usage: tools/gpu_r2_evalbwd.sh <outdir-name>: the eval-mode (frozen BatchNorm) backward test alone, then the whole GPU
# suite and smoke() on the same build
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 120 python -m pytest tests/test_gpu_tracknet.py -q -m gpu -s -k "eval_mode_backward" > $OUT/pytest_evalbwd.log 2>&1; echo "evalbwd rc=$?" > $OUT/summary.txt
grep -E "passed|failed|Error|assert|gradcheck" $OUT/pytest_evalbwd.log | tail -12 >> $OUT/summary.txt
timeout -k 5 300 python -m pytest tests -q -m gpu --timeout=200 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/summary.txt
tail -3 $OUT/pytest_gpu.log >> $OUT/summary.txt
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary.txt
tail -1 $OUT/smoke.log >> $OUT/summary.txt
cat $OUT/summary.txt
