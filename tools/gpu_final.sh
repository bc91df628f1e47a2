#!/bin/bash
# usage: tools/gpu_final.sh <outdir-name>: tests, bench (both precisions), configs[3]/[4], then the profiling pass
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
timeout -k 5 900 python -m pytest tests -q -m gpu --timeout=600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $OUT/summary0.txt
tail -2 $OUT/pytest_gpu.log >> $OUT/summary0.txt
timeout -k 5 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/summary0.txt
timeout -k 5 600 python bench.py --steps 20 --warmup 3 --precision tf32like --no-cpu-baseline > $OUT/bench_tf32like.log 2>&1; echo "bench tf32like rc=$?" >> $OUT/summary0.txt
for f in bench bench_tf32like; do tail -1 $OUT/$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$f value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3), 'clocks', d['clocks'])
    for k,v in d['kernel_breakdown'].items(): print('   ',k, round(v['ms_per_step'],3), v['tflops'] and round(v['tflops'],1))
except Exception as e: print('parse failed',e)" >> $OUT/summary0.txt; done
timeout -k 5 600 python tools/bench_configs.py > $OUT/configs.log 2>&1; echo "configs rc=$?" >> $OUT/summary0.txt
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary0.txt
bash tools/gpu_profile.sh $1 > $OUT/profile.log 2>&1
cat $OUT/summary0.txt $OUT/summary.txt; cut -c1-600 $OUT/configs.log
