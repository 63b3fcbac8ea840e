#!/bin/bash
# gpu_retry.sh <timeout_s> <command...>: gpurun, retried while the pod answers "transient / busy" (nothing charged).
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient\|nothing was charged" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
cat /tmp/gpurun_last.log
