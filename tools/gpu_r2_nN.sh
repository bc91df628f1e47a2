#!/bin/bash
# usage (gpurun --gpus N): tools/gpu_r2_nN.sh <outdir-name> <N>: the driver's torchrun launch line on N GPUs of one box with
# the gradient allreduce overlapped with the encoder's backward (default) and as one call after it
# (TNB_ALLREDUCE_OVERLAP=0), then the one-GPU line of the same box
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
N=$2
mkdir -p $OUT
: > $OUT/summary.txt
P=29631
for ov in 1 0; do
  P=$((P+1))
  TNB_ALLREDUCE_OVERLAP=$ov timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_n${N}_ov$ov.log 2>&1
  tail -1 $OUT/bench_n${N}_ov$ov.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('N=$N overlap=$ov: value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'e2e ms',round(d['e2e']['ms_per_step'],3),'train_step ms',round(d['train_step']['ms_per_step'],3),d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
done
timeout -k 5 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-alt-precision > $OUT/bench_n1.log 2>&1
tail -1 $OUT/bench_n1.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('N=1: value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'e2e ms',round(d['e2e']['ms_per_step'],3),d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
cat $OUT/summary.txt
