#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout-seconds> [--gpus N] -- '<command>'   (retries while the pod answers "busy")
log=$1; shift; tmo=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "${extra[@]}" -- "$1" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "exit=$rc" >> "$log"; exit $rc; fi
  sleep 150
done
echo "exit=gave-up" >> "$log"
