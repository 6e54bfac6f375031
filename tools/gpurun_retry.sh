#!/bin/bash
# usage: gpurun_retry.sh <timeout-seconds> '<command>'   (retries while the pod answers busy; nothing is charged for those)
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" ${GPUS:+--gpus $GPUS} -- "$2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 75; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
