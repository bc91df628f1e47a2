#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
for m in 0 1; do
TNB_CONV_MERGE=$m timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --per-launch > $OUT/bench_merge$m.log 2> $OUT/launches_merge$m.txt
done
paste -d'|' <(grep "^launch" $OUT/launches_merge0.txt | cut -c1-100) <(grep "^launch" $OUT/launches_merge1.txt | awk '{print $8, $10, $11, $12, $13}' ) | grep -E -- "->64:"
for m in 0 1; do tail -1 $OUT/bench_merge$m.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench merge=$m: ms',round(d['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks'])"; done
