#!/bin/bash
# tools/gpurun_retry_n.sh <gpus> <log> <timeout> <command...>: gpurun --gpus N with retries while the pod answers busy
n=$1; shift; log=$1; shift; to=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15; do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|no box\|status=busy" $log || [ $rc -eq 3 ]; then sleep 150; continue; fi
  break
done
exit $rc
