#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy" (exit code 3 / transient): tools/gpurun_retry.sh <log> <timeout> <command>
log=$1; to=$2; shift 2
for attempt in $(seq 1 20); do
    /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
    rc=$?
    if ! grep -q "status=transient" "$log" && [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
