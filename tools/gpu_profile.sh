#!/bin/bash
# Profiling pass for profiles/: launch list of one steady train step, tensor/DRAM metrics of every tensor-core launch,
# and two full captures. usage: tools/gpu_profile.sh <outdir-name>
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
export TNB_GRAPHS=0   # plain stream launches under the profiler (the library would otherwise replay CUDA graphs)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py 1 1 > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?" > $OUT/summary.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum
ncu --metrics $M --clock-control none -k regex:"conv3x3_kernel|wgrad3x3|view_presplit|bn_bwd_kernel" -s 94 --csv --log-file $OUT/tensor_metrics.csv python tools/profile_step.py 1 1 > $OUT/ncu_metrics.log 2>&1; echo "metrics rc=$?" >> $OUT/summary.txt
ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s 10 -c 1 -o $OUT/prof_conv_u1c1 python tools/profile_step.py 1 0 > $OUT/ncu_full1.log 2>&1; echo "full conv rc=$?" >> $OUT/summary.txt
ncu --set full --clock-control none --import-source on -k regex:wgrad3x3 -s 5 -c 1 -o $OUT/prof_wgrad_256 python tools/profile_step.py 1 0 > $OUT/ncu_full2.log 2>&1; echo "full wgrad rc=$?" >> $OUT/summary.txt
ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s 22 -c 1 -o $OUT/prof_dgrad_256 python tools/profile_step.py 1 0 > $OUT/ncu_full3.log 2>&1; echo "full dgrad rc=$?" >> $OUT/summary.txt
ncu --set full --clock-control none --import-source on -k regex:wgrad3x3_stacked -s 0 -c 1 -o $OUT/prof_wgrad_stacked python tools/profile_step.py 1 0 > $OUT/ncu_full4.log 2>&1; echo "full stacked wgrad rc=$?" >> $OUT/summary.txt
ncu --set full --clock-control none --import-source on -k regex:bn_bwd_kernel -s 1 -c 1 -o $OUT/prof_bn_bwd_apply python tools/profile_step.py 1 0 > $OUT/ncu_full5.log 2>&1; echo "full bn_bwd rc=$?" >> $OUT/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/nvsmi.csv 2>&1
cat $OUT/summary.txt
