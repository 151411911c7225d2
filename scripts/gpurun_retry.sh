#!/bin/bash
# dev helper (CPU container only): call gpurun, retrying while the pod answers "no box / slot right now" (exit 3).
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD"; fi
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
