#!/bin/bash
# Retry a gpurun call while the pod answers "no slot free" (exit code 3).  Usage: gpurun_retry.sh <tries> <gpurun args...>
tries=$1; shift
# the snapshot ships the built library: make sure it matches the sources (a stale one is rebuilt on the box, on GPU time)
python -c "import sys; sys.path.insert(0, '$(dirname "$0")/..'); from dfot_b200.build import build; build()" || exit 1
for i in $(seq 1 "$tries"); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
