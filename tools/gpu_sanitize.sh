#!/bin/bash
# usage: tools/gpu_sanitize.sh <outdir-name>: compute-sanitizer memcheck over the small-shape kernel tests
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
export TNB_GRAPHS=0
K="wgrad_tap_stacked or half_resolution or dgrad_fused or fused_views or temporal_ensemble_vs or frame_preprocessing or inpaintnet_backward or evaluate_on_gpu or small_forward or bn_relu_bwd"
timeout -k 5 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests -q -m gpu -x -k "$K" > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?" > $OUT/summary.txt
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/memcheck.log | head -20 >> $OUT/summary.txt
cat $OUT/summary.txt
