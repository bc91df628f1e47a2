#!/bin/bash
# First-contact diagnostics on the GPU box: every stage in its own process with its own timeout so that a
# trapping / hanging kernel cannot take the rest of the run with it. Logs land in gpurun_out/diag/.
cd "$(dirname "$0")/.."
OUT=gpurun_out/diag
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
for v in 0 1 2 3; do
  timeout -k 5 180 python tools/gpu_probe.py conv $v > $OUT/probe_conv_$v.log 2>&1; echo "probe conv $v rc=$?" >> $OUT/summary.txt
done
for v in 0 1 2 3; do
  timeout -k 5 180 python tools/gpu_probe.py wgrad $v > $OUT/probe_wgrad_$v.log 2>&1; echo "probe wgrad $v rc=$?" >> $OUT/summary.txt
done
for f in test_gpu_ops test_gpu_conv test_gpu_tracknet; do
  timeout -k 5 900 python -m pytest tests/$f.py -q -m gpu --timeout=300 -s > $OUT/$f.log 2>&1; echo "$f rc=$?" >> $OUT/summary.txt
  tail -5 $OUT/$f.log >> $OUT/summary.txt
done
cat $OUT/summary.txt
tail -30 $OUT/probe_conv_0.log
tail -12 $OUT/probe_wgrad_0.log
