#!/bin/bash
# HBM-bound kernels: tests, then the micro-benchmarks (norm / sampler / U-ViT glue).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; echo "=== $name"; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "rc=$?"; tail -n ${TAILN:-14} "gpurun_out/$name.log"; }
TAILN=4 run u_norm 300 python -m pytest tests/test_gpu_uvit_kernels.py -q --timeout 120
TAILN=4 run u_fwd 600 python -m pytest tests/test_gpu_parity.py -q -k "uvit" --timeout 300
run b_norm 300 python scripts/bench_kernels.py norm
run b_sampler 300 python scripts/bench_kernels.py sampler
run b_uvit 300 python scripts/bench_kernels.py uvit
TAILN=3 run fwd 300 python scripts/profile_forward.py 4 3
