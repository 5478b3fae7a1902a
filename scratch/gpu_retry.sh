#!/bin/bash
# usage: gpu_retry.sh <timeout> <script> [extra gpurun args]
for k in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $1 ${@:3} -- "bash $2" 2>&1)
  if echo "$out" | grep -q "status=transient\|rc=3\|rc=2 "; then sleep 90; continue; fi
  echo "$out" | tail -40
  exit 0
done
echo "gave up"
