#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / "transient": nothing is charged).
# Usage: tools/gpurun_retry.sh [gpurun options] -- '<command>'
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] no slot (attempt $i); sleeping 90 s"
  sleep 90
done
exit 3
