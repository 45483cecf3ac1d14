#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; echo "=== $name"; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "rc=$?"; tail -n ${TAILN:-4} "gpurun_out/$name.log"; }
run t_gemm 300 python -m pytest tests/test_gpu_kernels.py -q -k "gemm" --timeout 120
run u_conv 300 python -m pytest tests/test_gpu_uvit_kernels.py -q -k "conv or gemm" --timeout 120
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600
TAILN=30 run b_gemm 300 python scripts/bench_kernels.py uvit_gemm
TAILN=12 run b_gemm_k600 300 python scripts/bench_kernels.py gemm
