#!/bin/bash
# usage (under `gpurun --gpus 4`): tools/gpu_n4.sh <outdir-name> [N]: the driver's N-GPU launch of bench.py only
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
N=${2:-4}
mkdir -p $OUT
timeout -k 5 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.log 2>&1; echo "bench n$N rc=$?" > $OUT/summary.txt
tail -1 $OUT/bench_n$N.log | cut -c1-400 >> $OUT/summary.txt
cat $OUT/summary.txt
