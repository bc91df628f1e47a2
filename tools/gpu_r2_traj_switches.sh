#!/bin/bash
# usage: tools/gpu_r2_traj_switches.sh <outdir-name>: the 20-step Adam trajectory (tools/diag_adam.py, first lines) under
# the launch-machinery switches - CUDA-graph replay off, programmatic dependent launch off, both off
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
run() { name=$1; shift; env "$@" DIAG_ADAM_SHORT=1 timeout 200 python tools/diag_adam.py > $OUT/traj_$name.txt 2>&1; echo "== $name"; grep -E "^ours|^oracle fp32 \+ torch" $OUT/traj_$name.txt; }
run default
run nographs TNB_GRAPHS=0
run nopdl TNB_PDL=0
run neither TNB_GRAPHS=0 TNB_PDL=0
