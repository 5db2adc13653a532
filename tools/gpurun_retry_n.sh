#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout-seconds> '<command>'
G=$1; T=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 90 s"
  sleep 90
done
exit 3
