#!/bin/bash
# usage: tools/gpu_r2_fp16bwd.sh <outdir-name>: parity suite with the backward pass on fp16 pairs of dz * 2^k (22-bit
# operands) - then, on the same box, the 20-step Adam trajectory and the bench with fp16 pairs vs bf16 pairs
# (TNB_BWD_FMT=bf16: 16-bit operands, the default)
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 900 python -m pytest tests -x -q -m gpu --timeout=600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $OUT/summary.txt
tail -14 $OUT/pytest_gpu.log | cut -c1-300 >> $OUT/summary.txt
for bf in fp16 bf16; do
  TNB_BWD_FMT=$bf timeout -k 5 300 python -m pytest tests/test_gpu_tracknet.py -x -q -m gpu -s -k "twenty_adam or c2_shape" > $OUT/traj_bf$bf.log 2>&1
  echo "trajectory + c2 yardstick TNB_BWD_FMT=$bf rc=$?" >> $OUT/summary.txt
  grep -E "step (0|4|9|14|19):|passed|failed" $OUT/traj_bf$bf.log | cut -c1-200 >> $OUT/summary.txt
done
run() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 5 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --per-launch > $OUT/bench_$name.log 2> $OUT/launches_$name.txt
  tail -1 $OUT/bench_$name.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('bench $name: ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'train_step ms',round(d['train_step']['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks']['sm_mhz'])
except Exception as e: print('bench $name: parse failed',e)" >> $OUT/summary.txt
}
run fp16 TNB_BWD_FMT=fp16
run bf16 TNB_BWD_FMT=bf16
run fp16b TNB_BWD_FMT=fp16
cat $OUT/summary.txt
