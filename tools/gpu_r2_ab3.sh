#!/bin/bash
# usage: tools/gpu_r2_ab3.sh <outdir-name>: parity suite on the current kernels, then per-launch A/B on ONE box of
# resident weights (TNB_CONV_RESIDENT), the merged 64-wide tiles (TNB_CONV_MERGE) and the tile plan (TNB_CONV_PLAN)
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 900 python -m pytest tests -x -q -m gpu --timeout=600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $OUT/summary.txt
tail -14 $OUT/pytest_gpu.log | cut -c1-300 >> $OUT/summary.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 5 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --per-launch > $OUT/bench_$name.log 2> $OUT/launches_$name.txt
  tail -1 $OUT/bench_$name.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('bench $name: ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'train_step ms',round(d['train_step']['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks']['sm_mhz'], 'adam', round(d['train_step']['adam_kernel']['ms'],4))
except Exception as e: print('bench $name: parse failed',e)" >> $OUT/summary.txt
}
run res1 TNB_CONV_RESIDENT=1
run res0 TNB_CONV_RESIDENT=0
run merge0 TNB_CONV_MERGE=0
run plan0 TNB_CONV_PLAN=0
run res1b TNB_CONV_RESIDENT=1
cat $OUT/summary.txt
paste -d'|' <(grep "^launch" $OUT/launches_res1.txt | cut -c1-64) <(grep "^launch" $OUT/launches_res0.txt | awk '{print $8}') <(grep "^launch" $OUT/launches_merge0.txt | awk '{print $8}') <(grep "^launch" $OUT/launches_plan0.txt | awk '{print $8}') | grep -E "fwd|dgrad" | grep -E " 64->| 32->|->64:"
