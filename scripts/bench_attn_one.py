#!/usr/bin/env python
"""One attention shape, a few launches (ncu target).  Usage: bench_attn_one.py R heads dh N [iters] [score_bound]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfot_b200 import ops  # noqa: E402

R, heads, dh, N = (int(a) for a in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 4
bound = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
D = heads * dh
qkv = (torch.randn((R * N, 3 * D), device="cuda") * 0.5).to(torch.bfloat16)
out = torch.empty((R * N, D), device="cuda", dtype=torch.bfloat16)
for _ in range(iters):
    ops.attention(qkv, out, R, N, heads, dh, score_bound=bound)
torch.cuda.synchronize()
print("ok")
