#!/bin/bash
# usage: tools/gpu_pair_conv.sh <outdir-name>: first GPU run of the experimental CTA-pair forward/dgrad kernel
# (conv_pair.cu, TNB_CONV_PAIR=1): conv parity tests, whole-network tests, then bench with and without it.
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
TNB_CONV_PAIR=1 timeout -k 5 200 python -m pytest tests/test_gpu_conv.py -x -q -k "not wgrad and not bn_reduce" > $OUT/pytest_conv_pair.log 2>&1; echo "pytest conv (pair) rc=$?" > $OUT/summary.txt
tail -15 $OUT/pytest_conv_pair.log | cut -c1-300 >> $OUT/summary.txt
TNB_CONV_PAIR=1 timeout -k 5 200 python -m pytest tests/test_gpu_tracknet.py -x -q > $OUT/pytest_net_pair.log 2>&1; echo "pytest tracknet (pair) rc=$?" >> $OUT/summary.txt
tail -5 $OUT/pytest_net_pair.log | cut -c1-300 >> $OUT/summary.txt
for pair in 0 1; do
TNB_CONV_PAIR=$pair timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --per-launch > $OUT/bench_pair$pair.log 2> $OUT/launches_pair$pair.txt
tail -1 $OUT/bench_pair$pair.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('bench pair=$pair: ms',round(d['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
done
cat $OUT/summary.txt
paste -d'|' <(grep "^launch" $OUT/launches_pair0.txt | cut -c1-100) <(grep "^launch" $OUT/launches_pair1.txt | awk '{print $8, $10, $11, $12, $13}' ) | grep -E "fwd|dgrad"
