#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/run7
mkdir -p $OUT
timeout -k 5 1200 python -m pytest tests -q -m gpu --timeout=600 -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $OUT/summary.txt
grep -E "passed|failed|FAILED|worst gradient|C1 heatmap|gradcheck" $OUT/pytest_gpu.log | tail -15 >> $OUT/summary.txt
timeout -k 5 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/summary.txt; tail -3 $OUT/bench.log >> $OUT/summary.txt
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py 1 1 > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $OUT/summary.txt
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s 17 -c 1 -o $OUT/prof_dgrad python tools/profile_step.py 1 0 > $OUT/ncu_dgrad.log 2>&1; echo "ncu dgrad rc=$?" >> $OUT/summary.txt
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:wgrad3x3_kernel -s 0 -c 1 -o $OUT/prof_wgrad python tools/profile_step.py 1 0 > $OUT/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
