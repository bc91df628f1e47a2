#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 900 python -m pytest tests -q -m gpu --timeout=600 -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $OUT/summary.txt
tail -5 $OUT/pytest_gpu.log >> $OUT/summary.txt
timeout -k 5 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/summary.txt
python - <<PY >> $OUT/summary.txt
import json
try:
    d=json.loads(open("$OUT/bench.log").read().strip().splitlines()[-1])
    print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "e2e ms", round(d["e2e"]["ms_per_step"],3))
    for k,v in d["kernel_breakdown"].items(): print("   ",k, round(v["ms_per_step"],3), v["tflops"] and round(v["tflops"],1))
except Exception as e: print("bench parse failed", e)
PY
timeout -k 5 600 python tools/torch_cuda_baseline.py 10 5 > $OUT/torch_cuda.log 2>&1; echo "torch-cuda rc=$?" >> $OUT/summary.txt
grep impl $OUT/torch_cuda.log | cut -c1-160 >> $OUT/summary.txt
timeout -k 5 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.log 2>&1; echo "ref arm rc=$?" >> $OUT/summary.txt
tail -1 $OUT/bench_ref.log | cut -c1-300 >> $OUT/summary.txt
cat $OUT/summary.txt
