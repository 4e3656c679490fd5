#!/bin/bash
# usage: tools/gpu_retry.sh <gpus> <timeout> <script>; retries while the pod answers "busy" (exit 3), at most ~40 min
for i in $(seq 1 14); do
  /usr/local/graft/bin/gpurun --gpus $1 --timeout $2 -- "bash $3" > gpurun_out/retry_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/retry_last.log; then break; fi
  sleep 120
done
tail -120 gpurun_out/retry_last.log
