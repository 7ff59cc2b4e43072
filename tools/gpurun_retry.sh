#!/bin/bash
# retries a gpurun call while the pod answers "busy" (exit 3: nothing charged). Usage: [GPUS=n] tools/gpurun_retry.sh <timeout> <command...>
T=$1; shift
G=${GPUS:+--gpus $GPUS}
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 45
done
exit 3
