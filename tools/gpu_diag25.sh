#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
{
python tools/diag_smoke.py 0
TNB_GRAPHS=0 python tools/diag_smoke.py 0
TNB_CONV_PLAN=0 python tools/diag_smoke.py 0
TNB_CPASYNC_CA=0 python tools/diag_smoke.py 0
python tools/diag_smoke.py 128
python tools/diag_smoke.py 32
python tools/diag_smoke.py 0 128 192
python tools/diag_smoke.py 0 96 160
} > $OUT/diag.log 2>&1
timeout 600 python -m pytest tests -q -m gpu -x -k "evaluate or adam" > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log >> $OUT/diag.log
cat $OUT/diag.log
