#!/bin/bash
# usage: tools/gpu_r2_check.sh <outdir-name>: GPU parity suite, the train.py / predict.py command lines, small-batch latency
# under the PDL modes, BASELINE configs[3] / [4]
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 900 python -m pytest tests -x -q -m gpu --timeout=600 --durations=6 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $OUT/summary.txt
tail -14 $OUT/pytest_gpu.log | cut -c1-300 >> $OUT/summary.txt
bash tools/gpu_entrypoints.sh $1/entry > /dev/null 2>&1
cat $OUT/entry/summary.txt >> $OUT/summary.txt
for m in 0 1 2 3; do TNB_PDL=$m timeout -k 5 120 python tools/bench_configs.py latency 2>/dev/null | tail -1 >> $OUT/summary.txt; done
TNB_PDL=0 TNB_GRAPHS=0 timeout -k 5 120 python tools/bench_configs.py latency 2>/dev/null | tail -1 >> $OUT/summary.txt
TNB_PDL=3 TNB_GRAPHS=0 timeout -k 5 120 python tools/bench_configs.py latency 2>/dev/null | tail -1 >> $OUT/summary.txt
TNB_PDL=0 timeout -k 5 300 python tools/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?" >> $OUT/summary.txt
cut -c1-700 $OUT/configs.jsonl >> $OUT/summary.txt
cat $OUT/summary.txt
