#!/bin/bash
# Pair-kernel tile width sweep over the RE10K / K600 GEMM and conv shapes (DFOT_GEMM_BN pins the width).
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
for bn in 0 128 192 256; do
  echo "=== BN=$bn"
  DFOT_GEMM_BN=$bn python scripts/bench_kernels.py uvit_gemm 2>&1 | grep -E "^(conv|gemm)" | tee gpurun_out/bn_${bn}_uvit.log
  DFOT_GEMM_BN=$bn python scripts/bench_kernels.py gemm 2>&1 | grep -E "^gemm" | tee gpurun_out/bn_${bn}_k600.log
done
