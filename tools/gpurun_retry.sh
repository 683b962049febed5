#!/bin/bash
# retry a gpurun call while the pod answers "no box / slot free" (exit code 3; nothing is charged for those)
# usage: tools/gpurun_retry.sh <log file> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
