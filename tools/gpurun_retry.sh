#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'  -- retries while the pod answers "busy" (exit 3), logs to gpurun_out/.retry.log
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > gpurun_out/.retry.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 90
done
tail -150 gpurun_out/.retry.log
exit $rc
