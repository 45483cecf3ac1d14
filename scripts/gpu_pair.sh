#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; echo "=== $name"; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "rc=$?"; tail -n ${TAILN:-4} "gpurun_out/$name.log"; }
export DFOT_GEMM_PAIR=1
run p_gemm 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" --timeout 60
run p_conv 300 python -m pytest tests/test_gpu_uvit_kernels.py -q -x --timeout 60
run p_parity 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300
TAILN=30 run p_bench 300 python scripts/bench_kernels.py uvit_gemm
export DFOT_GEMM_PAIR=0
TAILN=30 run s_bench 300 python scripts/bench_kernels.py uvit_gemm
