#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -m pytest tests/test_gpu_uvit_kernels.py -q -k "side_output or conv3x3" --timeout 120 2>&1 | tail -2
python - <<'PY'
import math, sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "scripts")
from dfot_b200 import ops
from bench_kernels import timeit
for (n,H,C) in [(64,128,128),(64,64,256)]:
    x = torch.randn((n,H,H,C), device="cuda").to(torch.bfloat16)
    w = (torch.randn((C,3,3,C), device="cuda")/math.sqrt(9*C)).to(torch.bfloat16)
    bias = torch.randn((C,), device="cuda")
    sums = torch.empty((n,32,3), dtype=torch.float64, device="cuda")
    o16 = torch.empty((n*H*H, C), device="cuda", dtype=torch.bfloat16); o32 = torch.randn((n*H*H, C), device="cuda")
    for name, fn in [("bf16", lambda: ops.conv3x3_bf16(x,w,o16,ops.EPI_BF16,bias=bias)),
                     ("bf16+gn", lambda: ops.conv3x3_bf16(x,w,o16,ops.EPI_BF16,bias=bias,gn_sums=sums)),
                     ("resid", lambda: ops.conv3x3_bf16(x,w,o32,ops.EPI_RESID_F32,bias=bias,resid=o32)),
                     ("resid+gn", lambda: ops.conv3x3_bf16(x,w,o32,ops.EPI_RESID_F32,bias=bias,resid=o32,gn_sums=sums))]:
        us = timeit(fn, 20)
        print(f"conv {H}x{H} C={C} {name:9s}: {us:8.1f} us  {2.0*n*H*H*C*9*C/us/1e6:7.1f} TFLOP/s")
PY
