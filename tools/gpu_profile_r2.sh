#!/bin/bash
# usage: tools/gpu_profile_r2.sh <outdir-name>: the round-2 ncu passes of one steady-state train step (bs 10, fp32x3):
# (1) launch list of every kernel, (2) tensor-pipe / DRAM / L2 metrics of every tensor-core, BatchNorm-backward and view
# kernel, (3) `--set full` captures of one launch of each hot kernel. tools/refresh_profile_r2.py turns the output into
# profiles/r2_final.md + profiles/kernel_metrics_r2.json.
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
export TNB_GRAPHS=0   # plain stream launches under the profiler (the library would otherwise replay CUDA graphs)
export TNB_PDL=0      # ncu serialises launches anyway; keep the kernels' own durations free of dependent-launch waits
timeout -k 5 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py 1 1 > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?" > $OUT/summary.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum
# 84 matching launches per step (17 fwd + 16 dgrad + 17 wgrad + 34 bn_bwd; no view pass is left): skip the warm-up step's
timeout -k 5 200 ncu --metrics $M --clock-control none -k regex:"conv3x3|wgrad3x3|view_presplit|bn_bwd_kernel" -s 84 --csv --log-file $OUT/tensor_metrics.csv python tools/profile_step.py 1 1 > $OUT/ncu_metrics.log 2>&1; echo "metrics rc=$?" >> $OUT/summary.txt
full() {  # name, kernel regex, skip
  timeout -k 5 100 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o $OUT/$1 python tools/profile_step.py 0 1 > $OUT/ncu_$1.log 2>&1; echo "full $1 rc=$?" >> $OUT/summary.txt
}
full prof_fwd_64_64 'conv3x3_lean_kernel' 1                 # forward launches in layer order: 0 = 27(32)->64, 1 = 64->64
full prof_fwd_768_256 'conv3x3_lean_kernel' 10              # decoder up_block_1.conv_1
full prof_fwd_192_64 'conv3x3_lean_kernel' 15               # decoder up_block_3.conv_1, the longest forward launch
full prof_dgrad_256_256 'conv3x3_kernel<' 3                 # generic-loop dgrad launches (N side > 64): 64->192, 128->128, 128->384, 256->256
full prof_wgrad_pair 'wgrad3x3_pair_kernel' 1
full prof_wgrad_stacked 'wgrad3x3_stacked_kernel' 0         # up_block_3.conv_2, 64 -> 64 at 288x512
full prof_bn_bwd_apply 'bn_bwd_kernel<true' 0               # apply pass of up_block_3.conv_2 (HBM-bound)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/nvsmi.csv 2>&1
cat $OUT/summary.txt; ls -la $OUT
