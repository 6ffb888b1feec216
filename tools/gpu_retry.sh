#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>'   -- retries gpurun while the pod answers busy (rc 3)
T=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
