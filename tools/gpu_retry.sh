#!/bin/bash
# usage: tools/gpu_retry.sh <tag> <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers busy)
tag=$1; shift; tmo=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
mkdir -p gpurun_out
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "${extra[@]}" -- "$@" > gpurun_out/call_${tag}.txt 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 60
done
tail -40 gpurun_out/call_${tag}.txt
exit $rc
