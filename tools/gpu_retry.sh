#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>'  -- retries gpurun while the pod answers busy (exit 3 / transient)
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient"; then sleep 60; continue; fi
  break
done
