#!/bin/bash
# gpurun that retries while the pod answers "busy / draining" (exit 3: nothing charged).
#   tools/gpurun_retry.sh <log> [gpurun options] -- '<command>'
LOG=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  echo "exit $rc (attempt $i)" >> "$LOG"
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
