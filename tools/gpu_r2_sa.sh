#!/bin/bash
# usage: tools/gpu_r2_sa.sh <outdir-name>: number of halo-tile (A operand) stages of the conv kernel: 2 (default) vs 3 / 4 (TNB_CONV_SA),
# per-launch A/B on one box
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
: > $OUT/summary.txt
for cap in 22 23 33 22 23 33 32; do
  TNB_CONV_SA=$cap timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-alt-precision --per-launch 2> $OUT/launches_cap$cap.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('SA $cap: ms',round(d['ms_per_step'],3),{k:round(x['ms_per_step'],3) for k,x in d['kernel_breakdown'].items()},'clk',d['clocks']['sm_mhz'])" >> $OUT/summary.txt
done
cat $OUT/summary.txt
paste -d'|' <(grep "^launch" $OUT/launches_cap22.txt | cut -c1-62) <(grep "^launch" $OUT/launches_cap23.txt | awk '{print $8}') <(grep "^launch" $OUT/launches_cap33.txt | awk '{print $8}') | grep -E "fwd|dgrad" | head -0
