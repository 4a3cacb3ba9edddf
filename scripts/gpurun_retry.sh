#!/bin/bash
# usage: gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers "busy" (exit 3 / transient)
log=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > "$log" 2>&1
  if grep -q "status=transient\|no box or slot\|status=busy" "$log"; then sleep 120; continue; fi
  break
done
echo finished >> "$log"
