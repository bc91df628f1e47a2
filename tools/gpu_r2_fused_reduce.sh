#!/bin/bash
# usage: tools/gpu_r2_fused_reduce.sh <outdir-name>: where does the time go with the BatchNorm-backward reduction fused into
# the dgrad epilogue (variant bit 64)? Per-kind sums said -0.44 ms, the step said 0 (profiles/exp_r2a_summary.txt): ncu
# launch lists of one step with and without the bit, then the bench A/B/A on the same box.
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
export TNB_GRAPHS=0 TNB_PDL=0
for v in 0 64; do
  timeout -k 5 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_v$v.csv python tools/profile_step.py 1 1 fp32x3 10 $v > $OUT/ncu_v$v.log 2>&1
  python tools/summarize_launches.py $OUT/launches_v$v.csv | head -45 > $OUT/launches_v$v.md
done
unset TNB_GRAPHS TNB_PDL
for v in 0 64 0 64; do
  timeout -k 5 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-alt-precision --variant $v 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('variant $v: ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),{k:round(x['ms_per_step'],3) for k,x in d['kernel_breakdown'].items()},'clk',d['clocks']['sm_mhz'])" >> $OUT/summary.txt
done
cat $OUT/summary.txt
paste -d'|' <(cut -c1-75 $OUT/launches_v0.md) <(cut -c1-75 $OUT/launches_v64.md) | head -45
