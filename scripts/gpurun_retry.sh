#!/bin/bash
# gpurun with retries while the pod has no free slot (status=transient / exit code 3): usage gpurun_retry.sh <timeout> <command>
T=$1; shift
for i in $(seq 1 20); do
  out=$(gpurun --timeout $T -- "$@" 2>&1); rc=$?
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|no box or slot"; then sleep 90; continue; fi
  exit $rc
done
