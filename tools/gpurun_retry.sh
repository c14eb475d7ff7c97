#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod answers "busy" (nothing is charged then)
T=$1; shift
G=${GPUS:-1}
for i in $(seq 1 60); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  else out=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" 2>&1); fi
  if echo "$out" | grep -q "status=transient"; then sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "gave up: pod busy"; exit 3
