#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3 or status=transient, nothing charged).
#   tools/gpurun_retry.sh [--gpus N] [--timeout S] -- 'command'
for attempt in $(seq 1 30); do
    out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
    rc=$?
    echo "$out"
    if ! echo "$out" | grep -q -e "status=transient" -e "already running" && [ $rc -ne 3 ]; then exit $rc; fi
    echo "[retry] attempt $attempt answered busy; sleeping 90 s" >&2
    sleep 90
done
exit 3
