#!/bin/bash
# usage: gpu_retry.sh <timeout> <logfile> <command...>: retries while the pod answers "no slot" (exit 3), up to 12 times
to=$1; log=$2; shift 2
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log; then sleep 150; continue; fi
  exit $rc
done
