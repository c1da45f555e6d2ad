#!/bin/bash
# retry wrapper around gpurun: the pod answers "transient" while it drains slots (nothing charged)
# usage: scratch/gpu.sh <timeout-seconds> <script>
for i in 1 2 3 4 5 6 7 8 9 10; do
  out=$(timeout 3300 /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" 2>&1)
  echo "$out" | tail -6
  if echo "$out" | grep -q "status=transient"; then sleep 100; continue; fi
  break
done
