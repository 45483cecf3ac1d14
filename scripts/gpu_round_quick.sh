#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; echo "=== $name"; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "rc=$?"; tail -n ${TAILN:-3} "gpurun_out/$name.log"; }
run t_kernels 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 120
run t_uvit_kernels 600 python -m pytest tests/test_gpu_uvit_kernels.py -q --timeout 120
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600
TAILN=1 run bench_re10k 900 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline
TAILN=1 run bench_k600 900 python bench.py --workload k600 --steps 3 --warmup 3 --skip-cpu-baseline
TAILN=12 run b_gemm_k600 300 python scripts/bench_kernels.py gemm
