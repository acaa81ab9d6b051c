#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   — retries while the pod answers "busy" (exit 3, nothing charged)
log=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc" >> "$log"; exit $rc; fi
  sleep 90
done
echo "gpurun gave up (busy)" >> "$log"
