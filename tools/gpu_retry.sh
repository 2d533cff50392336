#!/bin/bash
# usage: tools/gpu_retry.sh TIMEOUT 'command'   -- retries while the pod answers "busy" (exit code 3)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 60
done
tail -25 /tmp/gpurun_last.log
exit $rc
