#!/bin/bash
# retry_gpurun.sh <timeout> <script> [--gpus N]: keeps asking for a box until the call is not refused as transient
T=$1; S=$2; shift 2
for i in $(seq 1 40); do
  out=$(gpurun --timeout $T "$@" -- "bash $S" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gave up"
