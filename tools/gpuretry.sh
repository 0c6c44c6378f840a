#!/bin/bash
# tools/gpuretry.sh <timeout-seconds> '<command>' — gpurun, retried while the pod answers "transient / busy" (nothing is charged for those)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpuretry.log 2>&1
  if grep -q "status=transient\|status=busy" /tmp/gpuretry.log || grep -q "exit code 3" /tmp/gpuretry.log; then sleep 120; continue; fi
  break
done
tail -40 /tmp/gpuretry.log
