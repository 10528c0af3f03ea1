#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> [--gpus N] -- '<command>'   (retries while the pod answers "busy")
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
