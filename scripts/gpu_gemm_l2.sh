#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/bench_one.py gemm_l2 5 > gpurun_out/plain_gemm_l2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/prof_gemm_l2 python scripts/bench_one.py gemm_l2 3 > gpurun_out/ncu_gemm_l2.log 2>&1
echo "rc=$?"; cat gpurun_out/plain_gemm_l2.log
python - <<'PY'
import math, sys, os, torch
sys.path.insert(0, os.getcwd())
from dfot_b200 import ops
sys.path.insert(0, "scripts")
from bench_kernels import timeit
for (M,N,K,epi) in [(65536,2304,576,ops.EPI_BF16),(65536,2304,576,ops.EPI_SILU_BF16),(65536,2304,1152,ops.EPI_BF16),(65536,2304,2304,ops.EPI_BF16),(65536,2304,288,ops.EPI_BF16),(65536,1728,576,ops.EPI_BF16),(65536,1792,576,ops.EPI_BF16)]:
    a = torch.randn((M, K), device="cuda").to(torch.bfloat16); w = (torch.randn((N, K), device="cuda")/math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn((N,), device="cuda"); out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    us = timeit(lambda: ops.gemm_bf16(a, w, out, epi, bias=bias), 20)
    print(f"M={M} N={N} K={K} epi={epi}: {us:8.1f} us {2.0*M*N*K/us/1e6:7.1f} TFLOP/s")
PY
