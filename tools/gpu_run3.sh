#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/run3
mkdir -p $OUT
timeout -k 5 1200 python -m pytest tests -q -m gpu --timeout=600 -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $OUT/summary.txt
grep -E "passed|failed|FAILED|worst gradient|C1 heatmap" $OUT/pytest_gpu.log | tail -15 >> $OUT/summary.txt
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary.txt; tail -2 $OUT/smoke.log >> $OUT/summary.txt
timeout -k 5 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/summary.txt; tail -3 $OUT/bench.log >> $OUT/summary.txt
timeout -k 5 300 python bench.py --steps 5 --warmup 3 --precision tf32like --no-cpu-baseline > $OUT/bench_tf32like.log 2>&1; echo "bench tf32like rc=$?" >> $OUT/summary.txt; tail -2 $OUT/bench_tf32like.log >> $OUT/summary.txt
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s 15 -c 1 -o $OUT/prof_conv python tools/profile_step.py 1 0 > $OUT/ncu_conv.log 2>&1; echo "ncu conv rc=$?" >> $OUT/summary.txt
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:wgrad3x3_kernel -s 1 -c 1 -o $OUT/prof_wgrad python tools/profile_step.py 1 0 > $OUT/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
