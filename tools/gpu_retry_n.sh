#!/bin/bash
# usage: tools/gpu_retry_n.sh <n-gpus> <timeout-seconds> '<command>'  -- like gpu_retry.sh on N GPUs of one box
N=$1; T=$2; shift 2
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  break
done
