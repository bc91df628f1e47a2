#!/bin/bash
# usage: tools/gpu_sanitize_r2.sh <outdir-name>: compute-sanitizer memcheck over the kernel-level GPU tests and the small
# network tests (plain stream launches; the full-size oracle comparisons are left out: they would run for an hour)
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
export TNB_GRAPHS=0
timeout -k 5 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest -q -m gpu -x \
  tests/test_gpu_conv.py tests/test_gpu_ops.py tests/test_gpu_tracknet.py tests/test_gpu_predict_flow.py \
  -k "not full_resolution and not full_size and not c2_shape and not bs10 and not twenty_adam and not resolution_sweep and not graph_replay and not c1_full and not checkpoint_interchanges and not is_the_derivative" \
  > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?" > $OUT/summary.txt
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/memcheck.log | head -20 >> $OUT/summary.txt
cat $OUT/summary.txt
