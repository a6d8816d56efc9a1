#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   — retries while the pod answers "busy / draining" (nothing charged)
log=$1; shift
for attempt in 1 2 3 4 5 6 7 8; do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    if grep -q "status=transient" "$log"; then sleep 120; continue; fi
    break
done
