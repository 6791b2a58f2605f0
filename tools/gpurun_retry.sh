#!/bin/bash
# tools/gpurun_retry.sh <log> <timeout> <command...>: gpurun with retries while the pod answers busy (exit 3 / transient)
log=$1; shift; to=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|no box\|status=busy" $log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
exit $rc
