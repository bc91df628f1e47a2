#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>': gpurun, retried while the pod answers "busy" (exit code 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
