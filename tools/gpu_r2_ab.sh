#!/bin/bash
# usage: tools/gpu_r2_ab.sh <outdir-name> [full]: GPU parity suite, then bench.py A/B of the round-2 switches on ONE box
# (TNB_PDL: programmatic dependent launch modes; TNB_CONV_LEAN: which MMA-issue loop), one complete default bench line
# (torch-CUDA baseline + CPU baseline included) and the backward-precision line (--precision fp32x3_bwd1)
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
: > $OUT/summary.txt
timeout -k 5 600 python -m pytest tests -x -q -m gpu --timeout=300 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/summary.txt
tail -15 $OUT/pytest_gpu.log | cut -c1-400 >> $OUT/summary.txt
run() {  # name, bench args, env...
  name=$1; shift; bargs=$1; shift
  env "$@" timeout -k 5 300 python bench.py --steps 20 --warmup 5 $bargs > $OUT/bench_$name.log 2> $OUT/launches_$name.txt
  tail -1 $OUT/bench_$name.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('bench $name: ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'train_step ms',round(d['train_step']['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks'],d['cuda_graphs'])
    if d.get('torch_cuda_baseline'): print('   torch_cuda_baseline', {k:(round(v['ms_per_step'],2) if isinstance(v,dict) and 'ms_per_step' in v else v) for k,v in d['torch_cuda_baseline'].items()})
    print('   adam', d['train_step']['adam_kernel'], 'mixup', d['train_step']['mixup_kernel'])
except Exception as e: print('bench $name: parse failed',e)" >> $OUT/summary.txt
}
Q="--no-cpu-baseline --no-torch-baseline --per-launch"
run pdl1 "$Q" TNB_PDL=1
run pdl0 "$Q" TNB_PDL=0
run pdl2 "$Q" TNB_PDL=2
run pdl3 "$Q" TNB_PDL=3
run pdl0b "$Q" TNB_PDL=0
run bwd1 "$Q --precision fp32x3_bwd1" TNB_PDL=0
run tf32like "$Q --precision tf32like" TNB_PDL=0
if [ "$2" = "full" ]; then run full "" X=1; fi
cat $OUT/summary.txt
