#!/bin/bash
# usage: tools/gpu_quick.sh <outdir-name> : gpu tests + bench + ncu launch list
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 1200 python -m pytest tests -q -m gpu --timeout=600 -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $OUT/summary.txt
grep -E "passed|failed|FAILED|worst gradient|C1 heatmap|gradcheck" $OUT/pytest_gpu.log | tail -15 >> $OUT/summary.txt
timeout -k 5 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/summary.txt
python - <<PY >> $OUT/summary.txt
import json
try:
    d=json.loads(open("$OUT/bench.log").read().strip().splitlines()[-1])
    print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
    for k,v in d["kernel_breakdown"].items(): print("   ",k, round(v["ms_per_step"],3), v["tflops"] and round(v["tflops"],1))
except Exception as e: print("bench parse failed", e)
PY
TNB_GRAPHS=0 timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py 1 1 > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
