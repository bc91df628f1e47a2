#!/bin/bash
# usage: tools/gpu_conv_experiments.sh <outdir-name>: first GPU run of the two experimental forward/dgrad kernels
# (conv_lean.cu: TNB_CONV_LEAN=1, conv_pair.cu: TNB_CONV_PAIR=1) - parity tests under each switch, then bench.py with the
# per-launch table for the shipped kernel and each experiment on the same box (plus TNB_BN_REVERSE=1: descending-order
# BatchNorm-backward reduction pass, and TNB_WGRAD_LEAN=1: lean issue loop of the tap-stacked wgrad kernel).
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
: > $OUT/summary.txt
TNB_BN_REVERSE=1 timeout -k 5 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tracknet.py -x -q -k "bn or train_step or backward" > $OUT/pytest_bnrev.log 2>&1; echo "pytest (TNB_BN_REVERSE) rc=$?" >> $OUT/summary.txt
tail -3 $OUT/pytest_bnrev.log | cut -c1-300 >> $OUT/summary.txt
TNB_WGRAD_LEAN=1 timeout -k 5 200 python -m pytest tests/test_gpu_conv.py -x -q -k "wgrad" > $OUT/pytest_wgrad_lean.log 2>&1; echo "pytest wgrad (TNB_WGRAD_LEAN) rc=$?" >> $OUT/summary.txt
tail -3 $OUT/pytest_wgrad_lean.log | cut -c1-300 >> $OUT/summary.txt
for sw in TNB_CONV_LEAN TNB_CONV_PAIR; do
  env $sw=1 timeout -k 5 200 python -m pytest tests/test_gpu_conv.py -x -q -k "not wgrad and not bn_reduce" > $OUT/pytest_conv_$sw.log 2>&1; echo "pytest conv ($sw) rc=$?" >> $OUT/summary.txt
  tail -12 $OUT/pytest_conv_$sw.log | cut -c1-300 >> $OUT/summary.txt
  env $sw=1 timeout -k 5 200 python -m pytest tests/test_gpu_tracknet.py -x -q > $OUT/pytest_net_$sw.log 2>&1; echo "pytest tracknet ($sw) rc=$?" >> $OUT/summary.txt
  tail -4 $OUT/pytest_net_$sw.log | cut -c1-300 >> $OUT/summary.txt
done
for sw in NONE TNB_CONV_LEAN TNB_CONV_PAIR TNB_BN_REVERSE TNB_WGRAD_LEAN; do
  env $sw=1 timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --per-launch > $OUT/bench_$sw.log 2> $OUT/launches_$sw.txt
  tail -1 $OUT/bench_$sw.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('bench $sw: ms',round(d['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
done
TNB_CONV_PAIR=1 TNB_CONV_LEAN=1 timeout -k 5 200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_tracknet.py -x -q -k "not wgrad and not bn_reduce" > $OUT/pytest_pair_lean.log 2>&1; echo "pytest (pair + lean) rc=$?" >> $OUT/summary.txt
tail -3 $OUT/pytest_pair_lean.log | cut -c1-300 >> $OUT/summary.txt
TNB_CONV_PAIR=1 TNB_CONV_LEAN=1 timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --per-launch > $OUT/bench_PAIR_LEAN.log 2> $OUT/launches_PAIR_LEAN.txt
tail -1 $OUT/bench_PAIR_LEAN.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('bench pair+lean: ms',round(d['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
TNB_CONV_LEAN=1 TNB_CONV_BACKOFF=1 timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_LEAN_BACKOFF.log 2>/dev/null
tail -1 $OUT/bench_LEAN_BACKOFF.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('bench lean + sleeping waits (TNB_CONV_BACKOFF): ms',round(d['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
TNB_CONV_LEAN=1 timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variant 64 > $OUT/bench_LEAN_FUSED.log 2>/dev/null
tail -1 $OUT/bench_LEAN_FUSED.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('bench lean + fused BN reduce (--variant 64): ms',round(d['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
TNB_CONV_LEAN=1 TNB_CONV_REMAP=1 timeout -k 5 200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_tracknet.py -x -q -k "not wgrad and not bn_reduce" > $OUT/pytest_remap.log 2>&1; echo "pytest (lean + remap) rc=$?" >> $OUT/summary.txt
tail -3 $OUT/pytest_remap.log | cut -c1-300 >> $OUT/summary.txt
TNB_CONV_LEAN=1 TNB_CONV_REMAP=1 timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_REMAP.log 2>/dev/null
tail -1 $OUT/bench_REMAP.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('bench lean+remap: ms',round(d['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
TNB_CONV_LEAN=1 TNB_PINGPONG=1 timeout -k 5 200 python -m pytest tests/test_gpu_tracknet.py -x -q > $OUT/pytest_pingpong.log 2>&1; echo "pytest tracknet (lean + pingpong) rc=$?" >> $OUT/summary.txt
tail -3 $OUT/pytest_pingpong.log | cut -c1-300 >> $OUT/summary.txt
TNB_CONV_LEAN=1 TNB_PINGPONG=1 timeout -k 5 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_PINGPONG.log 2>/dev/null
tail -1 $OUT/bench_PINGPONG.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('bench lean+pingpong: ms',round(d['ms_per_step'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()},'clk',d['clocks'])
except Exception as e: print('parse failed',e)" >> $OUT/summary.txt
cat $OUT/summary.txt
paste -d'|' <(grep "^launch" $OUT/launches_NONE.txt | cut -c1-100) <(grep "^launch" $OUT/launches_TNB_CONV_LEAN.txt | awk '{print $8}') <(grep "^launch" $OUT/launches_TNB_CONV_PAIR.txt | awk '{print $8}') | grep -E "fwd|dgrad"
