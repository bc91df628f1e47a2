#!/bin/bash
# Short profiling pass (tight GPU budget): launch list, tensor/DRAM metrics, full captures of the decoder conv and of the
# CTA-pair wgrad kernel. usage: tools/gpu_profile2.sh <outdir-name>
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
export TNB_GRAPHS=0   # plain stream launches under the profiler (the library would otherwise replay CUDA graphs)
timeout -k 5 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py 1 1 > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?" > $OUT/summary.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum
timeout -k 5 150 ncu --metrics $M --clock-control none -k regex:"conv3x3_kernel|wgrad3x3|view_presplit|bn_bwd_kernel" -s 94 --csv --log-file $OUT/tensor_metrics.csv python tools/profile_step.py 1 1 > $OUT/ncu_metrics.log 2>&1; echo "metrics rc=$?" >> $OUT/summary.txt
timeout -k 5 80 ncu --set full --clock-control none --import-source on -k regex:wgrad3x3_pair -s 1 -c 1 -o $OUT/prof_wgrad_256 python tools/profile_step.py 0 1 > $OUT/ncu_full2.log 2>&1; echo "full wgrad pair rc=$?" >> $OUT/summary.txt
timeout -k 5 80 ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s 10 -c 1 -o $OUT/prof_conv_u1c1 python tools/profile_step.py 0 1 > $OUT/ncu_full1.log 2>&1; echo "full conv rc=$?" >> $OUT/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/nvsmi.csv 2>&1
cat $OUT/summary.txt; ls -la $OUT
