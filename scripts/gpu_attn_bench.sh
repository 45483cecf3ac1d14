#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; echo "=== $name"; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "rc=$?"; tail -n ${TAILN:-6} "gpurun_out/$name.log"; }
run t_attn 300 python -m pytest tests/test_gpu_kernels.py -q -k "attention" --timeout 120
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q --timeout 300
TAILN=2 run bench 900 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline
