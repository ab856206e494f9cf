#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <command>   -- retries while gpurun reports "no box / slot free" (nothing charged)
T=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
tail -120 /tmp/gpurun_last.log
echo "gpurun rc=$rc attempts=$i"
