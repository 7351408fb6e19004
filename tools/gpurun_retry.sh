#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit 3).  usage: tools/gpurun_retry.sh <tries> <timeout-s> '<command>'
TRIES=$1; LIMIT=$2; shift 2
for i in $(seq 1 $TRIES); do
  /usr/local/graft/bin/gpurun --timeout $LIMIT -- "$@"; rc=$?
  echo "[retry] attempt $i rc=$rc"
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
