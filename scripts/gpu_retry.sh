#!/bin/bash
# gpurun wrapper: retry while the pod answers "transient" (queue timed out, nothing charged).  usage: gpu_retry.sh TIMEOUT 'command'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2" > gpurun_out/.retry.out 2>&1
  if ! grep -q "status=transient" gpurun_out/.retry.out; then break; fi
  echo "try $i: transient" >> gpurun_out/.retry.log
  sleep 60
done
tail -90 gpurun_out/.retry.out
