#!/bin/bash
# Development aid: gpurun with retries while the pod has no free slot.  Usage: scripts/grun.sh <log> <timeout> [--gpus N] -- '<command>'
log=$1; to=$2; shift 2
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
    /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
    if grep -q "status=transient\|status=busy\|exit code 3" $log || [ $? -eq 3 ]; then sleep 45; continue; fi
    break
done
