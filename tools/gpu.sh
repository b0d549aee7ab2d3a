#!/bin/bash
# usage: tools/gpu.sh [--gpus N] <timeout_s> '<command>'   -- gpurun with retries while the pod is busy (exit code 3)
extra=()
if [ "$1" = "--gpus" ]; then extra=(--gpus "$2"); shift 2; fi
t=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "${extra[@]}" --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
