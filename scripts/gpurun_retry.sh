#!/bin/bash
# usage: gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "busy / transient"
LOG=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=ok\|status=failed\|status=timeout" "$LOG"; then exit 0; fi
  sleep 150
done
