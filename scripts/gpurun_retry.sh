#!/bin/bash
# gpurun_retry.sh TIMEOUT 'command' : call gpurun until it gets a box (exit codes 2 / 3 / "transient" mean nothing ran).
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 120; continue; fi
  break
done
