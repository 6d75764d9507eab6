#!/bin/bash
# usage: scripts_gpurun_retry.sh <timeout> <script> [gpus]  -- retries while the pod answers "busy" (rc 3)
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "bash $S" > gpurun_out/call.log 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" > gpurun_out/call.log 2>&1; fi
  rc=$?
  if grep -q "status=transient" gpurun_out/call.log; then sleep 120; continue; fi
  break
done
tail -40 gpurun_out/call.log
