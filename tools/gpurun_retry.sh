#!/bin/bash
# Developer tool: gpurun, retried while the pod answers "busy" (exit code 3 / status=transient; nothing is charged for those).
#   tools/gpurun_retry.sh <log file> <gpurun arguments...>
log=$1; shift
for attempt in $(seq 1 40); do
	/usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
	if ! grep -q "status=transient" "$log"; then exit 0; fi
	sleep 90
done
exit 3
