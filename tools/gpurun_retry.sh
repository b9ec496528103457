#!/bin/bash
# gpurun wrapper: retry while the pod answers "busy" (exit code 3; nothing is charged for those).
# usage: tools/gpurun_retry.sh [gpurun flags] -- '<command>'
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
