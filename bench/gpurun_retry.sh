#!/bin/bash
# usage: bench/gpurun_retry.sh <logfile> <gpurun args...>   — retries while the pod answers "transient"/busy (nothing charged)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
