#!/usr/bin/env python
"""One kernel at one RE10K-size shape, a few launches (ncu target).
Usage: bench_one.py conv|gemm|gn_silu|sampler [iters]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfot_b200 import ops  # noqa: E402

which = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
DEV = "cuda"
if which == "conv":      # level-0 ResBlock second conv: 64 images 128x128, 128 -> 128 channels, + residual, f32 out
    n, H, C = 64, 128, 128
    x = torch.randn((n, H, H, C), device=DEV).to(torch.bfloat16)
    w = (torch.randn((C, 3, 3, C), device=DEV) / math.sqrt(9 * C)).to(torch.bfloat16)
    bias = torch.randn((C,), device=DEV)
    out = torch.randn((n * H * H, C), device=DEV)
    fn = lambda: ops.conv3x3_bf16(x, w, out, ops.EPI_RESID_F32, bias=bias, resid=out)
elif which == "gemm":    # level-3 fused qkv projection: M = 16384 tokens, 1152 -> 3456
    M, N, K = 16384, 3456, 1152
    a = torch.randn((M, K), device=DEV).to(torch.bfloat16)
    w = (torch.randn((N, K), device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn((N,), device=DEV)
    out = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    fn = lambda: ops.gemm_bf16(a, w, out, ops.EPI_BF16, bias=bias)
elif which == "gemm_l2":  # level-2 mlp half of the fused projection: M = 65536 tokens, 576 -> 2304, SiLU, bf16 out
    M, N, K = 65536, 2304, 576
    a = torch.randn((M, K), device=DEV).to(torch.bfloat16)
    w = (torch.randn((N, K), device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn((N,), device=DEV)
    out = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    fn = lambda: ops.gemm_bf16(a, w, out, ops.EPI_SILU_BF16, bias=bias)
elif which == "gn_silu":  # level-0 GroupNorm + FiLM (per-pixel pose part on half of the rows) + SiLU
    n, HW, C = 64, 16384, 128
    x = torch.randn((n * HW, C), device=DEV).to(torch.bfloat16)
    out = torch.empty((n * HW, C), device=DEV, dtype=torch.bfloat16)
    gamma, beta = torch.randn((C,), device=DEV), torch.randn((C,), device=DEV)
    sums = torch.empty((n, 32, 3), dtype=torch.float64, device=DEV)
    mod_img = torch.randn((n, 4 * C), device=DEV)
    mod_pix = torch.randn((32 * HW, 2 * C), device=DEV).to(torch.bfloat16)
    img_map = torch.tensor([(-1 if i % 16 < 8 else (i // 16) * 8 + i % 8) for i in range(n)], dtype=torch.int32, device=DEV)
    ops.groupnorm_stats(x, sums, n, HW, C)
    fn = lambda: ops.groupnorm_silu_bf16(x, sums, gamma, beta, out, n, HW, C, mod_img=mod_img, scale_col=0, shift_col=C,
                                         mod_pix=mod_pix, img_map=img_map)
elif which == "sampler":  # K4 at the RE10K shape: B=4, nfe=2, T=8, 3x256x256
    import numpy as np
    from dfot_b200.algorithms.dfot import sampling_plan as sp
    B, nfe, T, F = 4, 2, 8, 3 * 256 * 256
    x = torch.randn((B, T, F), device=DEV)
    mo = torch.randn((B * nfe, T, F), device=DEV)
    mi = torch.empty((B * nfe, T, F), device=DEV, dtype=torch.bfloat16)
    nh = torch.randn((B, T, F), device=DEV)
    upd = np.zeros((B * nfe, T), dtype=sp.UPDATE_DTYPE)
    upd["a"], upd["b"], upd["w"], upd["generate"] = 0.9, 0.1, 1.0, 1
    upd["w"][0::2] = -3.0
    upd["w"][1::2] = 4.0
    upd["generate"][:, 0] = 0
    prep = np.zeros((B * nfe, T), dtype=sp.PREPARE_DTYPE)
    prep["mode"][0::2, 0] = 1
    prep["qa"], prep["qb"] = 0.3, 0.9
    prep["noise_row"] = np.arange(B * nfe)[:, None] // nfe
    ud, pd = sp.to_device_bytes(upd, DEV), sp.to_device_bytes(prep, DEV)
    fn = lambda: ops.sampler_step_hg(x, mo, mi, ud, pd, None, nh, None, B, nfe, T)
elif which == "rmsnorm":  # level-2 RMSNorm + FiLM with the per-pixel pose part on half of the images: M = 65536, D = 576
    g, D = 32, 576
    M = 64 * g * g
    x = torch.randn((M, D), device=DEV)
    w = torch.randn((D,), device=DEV)
    mi = torch.randn((64, 4 * D), device=DEV)
    mp = torch.randn((32 * g * g, 2 * D), device=DEV).to(torch.bfloat16)
    o = torch.empty((M, D), device=DEV, dtype=torch.bfloat16)
    img_map = torch.tensor([(-1 if i % 16 < 8 else (i // 16) * 8 + i % 8) for i in range(64)], dtype=torch.int32, device=DEV)
    fn = lambda: ops.rmsnorm_film_bf16(x, w, mi, 0, D, g * g, o, mod_pix=mp, img_map=img_map)
elif which == "qknorm":   # level-2 q/k RMSNorm(head_dim 64) + RoPE-3D in place: M = 65536 tokens, 9 heads
    heads, dh, T, g = 9, 64, 8, 32
    D, Ntok = heads * dh, T * g * g
    M = 8 * Ntok
    qkv = torch.randn((M, 3 * D), device=DEV).to(torch.bfloat16)
    qw, kw = torch.randn((dh,), device=DEV), torch.randn((dh,), device=DEV)
    table = torch.randn((Ntok, dh // 2, 2), device=DEV)
    fn = lambda: ops.qk_norm_rope(qkv, qw, kw, table, Ntok, heads, dh, 0.18)
elif which == "rmsnorm1152":  # level-3 RMSNorm + FiLM: M = 16384 tokens, D = 1152 (pose part on half of the images)
    g, D = 16, 1152
    M = 64 * g * g
    x = torch.randn((M, D), device=DEV)
    w = torch.randn((D,), device=DEV)
    mi = torch.randn((64, 4 * D), device=DEV)
    mp = torch.randn((32 * g * g, 2 * D), device=DEV).to(torch.bfloat16)
    o = torch.empty((M, D), device=DEV, dtype=torch.bfloat16)
    img_map = torch.tensor([(-1 if i % 16 < 8 else (i // 16) * 8 + i % 8) for i in range(64)], dtype=torch.int32, device=DEV)
    fn = lambda: ops.rmsnorm_film_bf16(x, w, mi, 0, D, g * g, o, mod_pix=mp, img_map=img_map)
elif which == "gn_stats":  # stand-alone GroupNorm statistics of a level-0 bf16 feature map: 64 images 128x128x128
    n, HW, C = 64, 16384, 128
    x = torch.randn((n * HW, C), device=DEV).to(torch.bfloat16)
    sums = torch.empty((n, 32, 3), dtype=torch.float64, device=DEV)
    fn = lambda: ops.groupnorm_stats(x, sums, n, HW, C)
elif which == "gn_stats_f32":  # ... of the fp32 residual stream (the first ResBlock of a level)
    n, HW, C = 64, 16384, 128
    x = torch.randn((n * HW, C), device=DEV)
    sums = torch.empty((n, 32, 3), dtype=torch.float64, device=DEV)
    fn = lambda: ops.groupnorm_stats(x, sums, n, HW, C)
elif which == "adaln":     # K600 adaLN-LayerNorm launch: 8 x 1280 tokens, D = 1152, bf16 output only
    M, D, tpf = 8 * 1280, 1152, 256
    x = torch.randn((M, D), device=DEV)
    mod = torch.randn((M // tpf, 6 * D), device=DEV)
    y = torch.empty((M, D), device=DEV, dtype=torch.bfloat16)
    fn = lambda: ops.adaln_layernorm(x, mod, 0, D, tpf, y_bf16=y)
elif which in ("patch_mix", "patch_expand"):   # matrix-attention u factors: 64 rows x 16 frames x 256 patches, D = 768
    R, L, P, Mc, D = 64, 16, 256, 1, 768
    yv = torch.randn((R * L * P, D), device=DEV)
    if which == "patch_mix":
        u = torch.randn((P, Mc), device=DEV)
        o = torch.empty((R * Mc * L, D), device=DEV, dtype=torch.bfloat16)
        fn = lambda: ops.patch_mix_bf16(yv, u, o, R, L, P, Mc)
    else:
        pu, z = torch.randn((Mc, P), device=DEV), torch.randn((R * Mc * L, D), device=DEV)
        mod, xo = torch.randn((R * L, 6 * D), device=DEV), torch.empty_like(yv)
        fn = lambda: ops.patch_expand_gate_resid(xo, yv, z, pu, None, mod[:, 2 * D:], 6 * D, R, L, P, Mc)
else:
    raise SystemExit(f"unknown kernel {which}")
for _ in range(2):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fn()
e1.record()
torch.cuda.synchronize()
print(f"{which}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us per launch")
