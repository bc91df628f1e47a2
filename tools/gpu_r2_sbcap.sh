#!/bin/bash
# usage: tools/gpu_r2_sbcap.sh <outdir-name>: depth of the weight ring of the narrow (BN <= 128, 3 taps per stage) conv
# tiles: 3 slots (default) vs 4 / 5 (TNB_CONV_SB_CAP), per-launch A/B on one box
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
: > $OUT/summary.txt
for cap in 3 5 4 3; do
  TNB_CONV_SB_CAP=$cap timeout -k 5 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-alt-precision --per-launch 2> $OUT/launches_cap$cap.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('SB cap $cap: ms',round(d['ms_per_step'],3),{k:round(x['ms_per_step'],3) for k,x in d['kernel_breakdown'].items()},'clk',d['clocks']['sm_mhz'])" >> $OUT/summary.txt
done
cat $OUT/summary.txt
paste -d'|' <(grep "^launch" $OUT/launches_cap3.txt | cut -c1-62) <(grep "^launch" $OUT/launches_cap4.txt | awk '{print $8}') <(grep "^launch" $OUT/launches_cap5.txt | awk '{print $8}') | grep -E "fwd|dgrad" | grep -E " 64->|32->|->64:|128->128|64->128|128->64"
