#!/bin/bash
# usage: tools/gpu_retry.sh TIMEOUT_SECONDS 'command'   -- retries gpurun while the pod answers "busy" (nothing is charged then)
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out" | tail -25; exit $rc
done
echo "gave up: pod busy"; exit 3
