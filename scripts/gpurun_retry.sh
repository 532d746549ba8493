#!/bin/bash
# developer aid: retry a gpurun call while the pod answers "busy" (exit code 3)
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
