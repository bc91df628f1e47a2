#!/bin/bash
# usage: tools/gpu_sanitize.sh <outdir-name>: compute-sanitizer memcheck over the GPU test suite (plain stream launches)
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
export TNB_GRAPHS=0
timeout -k 5 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests -q -m gpu \
  --deselect tests/test_gpu_tracknet.py::test_cuda_graph_replay_equals_eager_launches > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?" > $OUT/summary.txt
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/memcheck.log | head -20 >> $OUT/summary.txt
cat $OUT/summary.txt
