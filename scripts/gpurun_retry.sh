#!/bin/bash
# usage: scripts/gpurun_retry.sh [gpurun options] -- 'command'   (retries while the pool answers "busy", exit code 3)
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
