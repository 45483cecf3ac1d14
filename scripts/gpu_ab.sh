#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for lib in "" "$PWD/diffusion-forcing-transformer_b200/build/lib_direct.so"; do
  echo "== lib=${lib:-default}"
  DFOT_B200_LIB=$lib python -m pytest tests/test_gpu_kernels.py -q -k "gemm" --timeout 120 2>&1 | tail -1
  DFOT_B200_LIB=$lib python scripts/bench_kernels.py uvit_gemm 2>&1 | grep "epi=1\|epi=3\|per forward"
done
