#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; echo "=== $name"; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "rc=$?"; tail -n ${TAILN:-14} "gpurun_out/$name.log"; }
run t_attn 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" --timeout 120
run b_attn2 300 python scripts/bench_kernels.py attn
true
