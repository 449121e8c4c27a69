#!/bin/bash
# tools/gpu_retry.sh TIMEOUT 'command' — gpurun with retries while the pod answers busy (exit 3: nothing charged)
T=${1:?timeout}; shift
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
