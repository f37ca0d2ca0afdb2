#!/bin/bash
# usage: scripts/gpurun_retry.sh [gpurun options] -- 'command'   - retries while the pod answers "busy" (exit code 3)
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 90 s" >&2
  sleep 90
done
exit 3
