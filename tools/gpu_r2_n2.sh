#!/bin/bash
# usage (gpurun --gpus 2): tools/gpu_r2_n2.sh <outdir-name>: the 2-GPU NCCL tests (averaged gradients == mean of the rank
# gradients, plain and overlapped), then the driver's torchrun launch line with the allreduce overlapped with the
# encoder's backward (default) and as one call after the backward (TNB_ALLREDUCE_OVERLAP=0), A/B/A on the same box,
# and the one-GPU line of the same box
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 400 python -m pytest tests/test_gpu_dp.py tests/test_gpu_tracknet.py -x -q -m gpu -k "nccl or two_ranges or data_parallel" --timeout=300 > $OUT/pytest_dp.log 2>&1; echo "pytest dp rc=$?" > $OUT/summary.txt
tail -4 $OUT/pytest_dp.log | cut -c1-300 >> $OUT/summary.txt
P=29531
for ov in 1 0; do
  P=$((P+1))
  TNB_ALLREDUCE_OVERLAP=$ov timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 30 --warmup 5 > $OUT/bench_n2_ov$ov.log 2>&1
  tail -1 $OUT/bench_n2_ov$ov.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('N=2 overlap=$ov: value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'e2e ms',round(d['e2e']['ms_per_step'],3),'train_step ms',round(d['train_step']['ms_per_step'],3),d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
done
timeout -k 5 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-alt-precision > $OUT/bench_n1.log 2>&1
tail -1 $OUT/bench_n1.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('N=1: value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'e2e ms',round(d['e2e']['ms_per_step'],3),d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
cat $OUT/summary.txt
