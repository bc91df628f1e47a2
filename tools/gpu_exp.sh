#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 100 python -m pytest tests/test_gpu_conv.py -x -q -k "wgrad" > $OUT/pytest.log 2>&1; echo "pytest wgrad rc=$?" > $OUT/summary.txt
tail -3 $OUT/pytest.log | cut -c1-300 >> $OUT/summary.txt
for cap in 3 5 3 5; do
TNB_CONV_SBCAP=$cap timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --per-launch > $OUT/bench_cap$cap.log 2> $OUT/launches_cap$cap.txt
tail -1 $OUT/bench_cap$cap.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench cap=$cap: ms',round(d['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks'])" >> $OUT/summary.txt
done
cat $OUT/summary.txt
paste -d'|' <(grep "^launch" $OUT/launches_cap3.txt | cut -c1-100) <(grep "^launch" $OUT/launches_cap5.txt | awk '{print $8, $10, $11, $12, $13}' ) | grep -E "fwd|dgrad" | grep -E -- "->64:|->128:"
