#!/bin/bash
# usage (under `gpurun --gpus 2`): tools/gpu_n2.sh <outdir-name> : the driver's N=2 launch of bench.py, then N=1 on the same box
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/bench_n2.log 2>&1; echo "bench n2 rc=$?" > $OUT/summary.txt
tail -1 $OUT/bench_n2.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('n2 value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'n_gpus',d['n_gpus'],d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
timeout -k 5 120 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_n1.log 2>&1; echo "bench n1 rc=$?" >> $OUT/summary.txt
tail -1 $OUT/bench_n1.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('n1 value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
cat $OUT/summary.txt
