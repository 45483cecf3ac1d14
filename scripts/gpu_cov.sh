#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; echo "=== $name"; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "rc=$?"; tail -n ${TAILN:-12} "gpurun_out/$name.log"; }
run t_cov 900 python -m pytest tests/test_gpu_parity.py -q -k "cfg5 or cfg4" --timeout 600
run smoke 600 python __graft_entry__.py smoke
