#!/bin/bash
# usage: tools/gpu_r2_u7.sh <outdir-name>: global loads in flight per producer thread of the forward gathers: 4 halo pixels
# (shipped) vs 7 (one batch per 18 x 18 tile): needs an alternate build with `constexpr int U0 ... : 7` in conv_kernel.inc linked as
# tracknetv3_b200/libtracknet_b200_u7.so (TNB_LIBRARY selects the library). Measured: forward 5.90 -> 5.81 ms, within the run-to-run spread
# (profiles/r2_u7_summary.txt): not adopted.
# per-launch A/B/A/B on one box
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
: > $OUT/summary.txt
ALT=$PWD/tracknetv3_b200/libtracknet_b200_u7.so
i=0
for lib in "" $ALT "" $ALT; do
  i=$((i+1))
  TNB_LIBRARY=$lib timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-alt-precision --per-launch 2> $OUT/launches_$i.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lib [$lib]: ms',round(d['ms_per_step'],3),{k:round(x['ms_per_step'],3) for k,x in d['kernel_breakdown'].items()},'clk',d['clocks']['sm_mhz'])" >> $OUT/summary.txt
done
cat $OUT/summary.txt
paste -d'|' <(grep "^launch" $OUT/launches_1.txt | cut -c1-52) <(grep "^launch" $OUT/launches_1.txt | awk '{print $8}') <(grep "^launch" $OUT/launches_2.txt | awk '{print $8}') <(grep "^launch" $OUT/launches_3.txt | awk '{print $8}') <(grep "^launch" $OUT/launches_4.txt | awk '{print $8}') | grep -E "fwd"
