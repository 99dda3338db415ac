#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  Usage: tools/gpurun_retry.sh [gpurun args] -- '<command>'
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $attempt answered busy; retrying in 90 s" >&2
  sleep 90
done
exit 3
