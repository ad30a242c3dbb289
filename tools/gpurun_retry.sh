#!/bin/bash
# Developer tool: gpurun, retried while the pod answers "busy" (status=transient; nothing is charged for those). An attempt only starts while
# /tmp/repo_busy does not exist (touch it while the tree is being rebuilt: the snapshot must not catch a library that does not match the sources).
#   tools/gpurun_retry.sh <log file> <gpurun arguments...>
log=$1; shift
for attempt in $(seq 1 80); do
	while [ -e /tmp/repo_busy ]; do sleep 5; done
	/usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
	if ! grep -q "status=transient" "$log"; then exit 0; fi
	sleep 45
done
exit 3
