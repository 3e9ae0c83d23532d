#!/usr/bin/env bash
# Development helper: gpurun answers 3 ("no box or slot free right now, nothing charged") on a busy pod — retry every 2
# minutes, up to 20 times. Usage: tools/gpurun_retry.sh [gpurun options] -- '<command>'
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  echo "[gpurun_retry] busy (attempt $attempt), sleeping 120 s" >&2
  sleep 120
done
exit 3
