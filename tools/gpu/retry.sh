#!/bin/bash
# usage: tools/gpu/retry.sh OUTFILE [gpurun args...]  -- retries while the pod answers "transient" (nothing charged)
out=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  if grep -q "status=transient" "$out"; then sleep 90; continue; fi
  break
done
