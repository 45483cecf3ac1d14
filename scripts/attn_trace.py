#!/usr/bin/env python
"""Phase timeline of attention kernel 2 (debug build: scripts/build_variant.sh trace attention_tcgen05.cu -DDFOT_ATTN_TRACE).

  DFOT_B200_LIB=.../variants/lib_trace.so python scripts/attn_trace.py R heads dh N [score_bound] > gpurun_out/trace.txt

Prints, for CTA 0, the clock64 stamps of every traced warp relative to the first stamp: softmax warps 4..11 (tags 1 S_FULL
acquired, 2 scores in registers, 3/4 before/after the PV_DONE wait, 5 P published) and the MMA issuers 1..2 (10 K ready,
11 S buffer free, 12 S issued, 13 V ready, 14 P ready, 15 PV issued)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfot_b200 import _abi, ops  # noqa: E402

R, heads, dh, N = (int(a) for a in sys.argv[1:5])
bound = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
D = heads * dh
qkv = (torch.randn((R * N, 3 * D), device="cuda") * 0.5).to(torch.bfloat16)
out = torch.empty((R * N, D), device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(qkv, out, R, N, heads, dh, score_bound=bound)
torch.cuda.synchronize()
L = _abi.lib()
buf = np.zeros(16 * 2048, dtype=np.uint64)
ops.attention(qkv, out, R, N, heads, dh, score_bound=bound)
rc = L.dfot_debug_attn_trace(ctypes.c_void_p(buf.ctypes.data), ctypes.c_int64(buf.nbytes))
assert rc == 0, rc
buf = buf.reshape(16, 2048)
t0 = min(int(b[0] >> 8) for b in buf if b[0])
for w in range(16):
    ev = [(int(x >> 8) - t0, int(x & 0xFF)) for x in buf[w] if x]
    if not ev:
        continue
    print(f"warp {w}: {len(ev)} events")
    print(" ".join(f"{t}:{tag}" for t, tag in ev[:int(os.environ.get('TRACE_EVENTS', '260'))]))
