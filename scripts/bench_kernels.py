#!/usr/bin/env python
"""Micro-benchmarks of the individual kernels at the K600 DiT-XL shapes (CUDA events on the launch stream).
Usage: python scripts/bench_kernels.py [attn|gemm|norm|all] [--iters N]"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfot_b200 import ops  # noqa: E402

DEV = "cuda"


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def bench_attn(iters):
    for (R, heads, dh, N) in [(8, 16, 72, 1280), (8, 16, 64, 1280), (8, 9, 128, 2048), (2, 12, 64, 576), (2, 9, 64, 8192), (8, 9, 64, 8192), (2, 9, 128, 2048)]:
        D = heads * dh
        qkv = (torch.randn((R * N, 3 * D), device=DEV) * 0.5).to(torch.bfloat16)
        out = torch.empty((R * N, D), device=DEV, dtype=torch.bfloat16)
        us = timeit(lambda: ops.attention(qkv, out, R, N, heads, dh), iters)
        fl = 4.0 * R * heads * N * N * dh
        print(f"attention R={R} heads={heads} d={dh} N={N}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
        us = timeit(lambda: ops.attention(qkv, out, R, N, heads, dh, score_bound=40.0), iters)
        print(f"  bounded scores (no running max):        {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")


def bench_gemm(iters):
    M = 10240
    for name, N, K, epi in [("qkv+rope", 3456, 1152, ops.EPI_QKV_ROPE_BF16), ("proj+gate", 1152, 1152, ops.EPI_GATE_RESID_F32),
                            ("fc1+gelu", 4608, 1152, ops.EPI_GELU_BF16), ("fc2+gate", 1152, 4608, ops.EPI_GATE_RESID_F32),
                            ("plain bf16 4608x1152", 4608, 1152, ops.EPI_BF16), ("plain f32 1152x1152", 1152, 1152, ops.EPI_F32),
                            ("plain bf16 1152x1152", 1152, 1152, ops.EPI_BF16), ("plain f32 1152x4608", 1152, 4608, ops.EPI_F32),
                            ("plain bf16 8192^3", 8192, 8192, ops.EPI_BF16)]:
        m = 8192 if "8192" in name else M
        a = torch.randn((m, K), device=DEV).to(torch.bfloat16)
        w = (torch.randn((N, K), device=DEV) / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn((N,), device=DEV)
        kw = dict(bias=bias)
        if epi == ops.EPI_GATE_RESID_F32:
            out = torch.empty((m, N), device=DEV)
            mod = torch.randn((m // 256, 3 * N), device=DEV)
            kw.update(resid=torch.randn((m, N), device=DEV), gate=mod[:, 2 * N:], ld_gate=3 * N, tokens_per_frame=256)
        elif epi == ops.EPI_QKV_ROPE_BF16:
            out = torch.empty((m, N), device=DEV, dtype=torch.bfloat16)
            kw.update(rope_cs=torch.randn((1280, 36, 2), device=DEV), tokens_per_sample=1280, model_dim=1152, head_dim=72,
                      q_scale=0.17)
        elif epi == ops.EPI_F32:
            out = torch.empty((m, N), device=DEV)
        else:
            out = torch.empty((m, N), device=DEV, dtype=torch.bfloat16)
        us = timeit(lambda: ops.gemm_bf16(a, w, out, epi, **kw), iters)
        print(f"gemm {name:22s} M={m} N={N} K={K}: {us:8.1f} us  {2.0 * m * N * K / us / 1e6:7.1f} TFLOP/s")
    a = torch.randn((8192, 8192), device=DEV).to(torch.bfloat16)
    us = timeit(lambda: torch.matmul(a, a.t()), iters)
    print(f"cuBLAS bf16 8192^3 (reference point): {us:8.1f} us  {2.0 * 8192 ** 3 / us / 1e6:7.1f} TFLOP/s")


def bench_norm(iters):
    for M, D, P in [(10240, 1152, 256), (81920, 1152, 256)]:      # K600 step (8 rows x 1280 tokens) and 8x that (asymptote)
        x = torch.randn((M, D), device=DEV)
        mod = torch.randn((M // P, 6 * D), device=DEV)
        y32, y16 = torch.empty((M, D), device=DEV), torch.empty((M, D), device=DEV, dtype=torch.bfloat16)
        us = timeit(lambda: ops.adaln_layernorm(x, mod, 0, D, P, y_f32=y32, y_bf16=y16), iters)
        print(f"adaln_layernorm M={M} D={D} (f32+bf16 out): {us:7.1f} us  {M * D * 10 / us / 1e3:7.1f} GB/s")
        us = timeit(lambda: ops.adaln_layernorm(x, mod, 0, D, P, y_bf16=y16), iters)
        print(f"adaln_layernorm M={M} D={D} (bf16 out):     {us:7.1f} us  {M * D * 6 / us / 1e3:7.1f} GB/s")


def bench_matrix(iters):
    """The u-factor kernels of the matrix-attention DiT variants (HBM-bound: one pass over the token stream each)."""
    for R, L, P, Mc, D in [(64, 16, 256, 1, 768), (16, 36, 16, 1, 512)]:
        y = torch.randn((R * L * P, D), device=DEV)
        u, pu = torch.randn((P, Mc), device=DEV), torch.randn((Mc, P), device=DEV)
        out = torch.empty((R * Mc * L, D), device=DEV, dtype=torch.bfloat16)
        z, mod = torch.randn((R * Mc * L, D), device=DEV), torch.randn((R * L, 6 * D), device=DEV)
        x = torch.empty_like(y)
        us = timeit(lambda: ops.patch_mix_bf16(y, u, out, R, L, P, Mc), iters)
        print(f"patch_mix R*L={R * L} P={P} D={D}:          {us:7.1f} us  {y.numel() * 4 / us / 1e3:7.1f} GB/s")
        us = timeit(lambda: ops.patch_expand_gate_resid(x, y, z, pu, None, mod[:, 2 * D:], 6 * D, R, L, P, Mc), iters)
        print(f"patch_expand_gate_resid (same shape):      {us:7.1f} us  {y.numel() * 8 / us / 1e3:7.1f} GB/s")


def bench_gemm_ditb(iters):
    """DiT-B (DMLab) block GEMMs: D = 768, 12 heads of 64, 16 tokens per frame, M = rows x frames x 16 tokens."""
    D, P = 768, 16
    for M in (36864, 147456):
        for name, N, K, epi in [("qkv+rope", 3 * D, D, ops.EPI_QKV_ROPE_BF16), ("qkv plain", 3 * D, D, ops.EPI_BF16),
                                ("proj+gate", D, D, ops.EPI_GATE_RESID_F32), ("proj plain f32", D, D, ops.EPI_F32),
                                ("fc1+gelu", 4 * D, D, ops.EPI_GELU_BF16), ("fc1 plain", 4 * D, D, ops.EPI_BF16),
                                ("fc2+gate", D, 4 * D, ops.EPI_GATE_RESID_F32), ("fc2 plain f32", D, 4 * D, ops.EPI_F32)]:
            a = torch.randn((M, K), device=DEV).to(torch.bfloat16)
            w = (torch.randn((N, K), device=DEV) / math.sqrt(K)).to(torch.bfloat16)
            kw = dict(bias=torch.randn((N,), device=DEV))
            if epi == ops.EPI_GATE_RESID_F32:
                out = torch.empty((M, N), device=DEV)
                mod = torch.randn((M // P, 3 * N), device=DEV)
                kw.update(resid=torch.randn((M, N), device=DEV), gate=mod[:, 2 * N:], ld_gate=3 * N, tokens_per_frame=P)
            elif epi == ops.EPI_QKV_ROPE_BF16:
                out = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
                kw.update(rope_cs=torch.randn((2304, 32, 2), device=DEV), tokens_per_sample=2304, model_dim=D, head_dim=64,
                          q_scale=0.18)
            elif epi == ops.EPI_F32:
                out = torch.empty((M, N), device=DEV)
            else:
                out = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
            us = timeit(lambda: ops.gemm_bf16(a, w, out, epi, **kw), iters)
            print(f"gemm {name:16s} M={M} N={N} K={K}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")


def bench_sampler(iters):
    """K4 (fused DDIM update + history-guidance combine + next-step inputs): algorithmic bytes per element =
    nfe*4 (model out f32) + 4 (x_t) + 4 (x_t+1) + nfe*2 (bf16 inputs) + 4 (history noise)."""
    import numpy as np
    from dfot_b200.algorithms.dfot import sampling_plan as sp
    for B, nfe, T, F in [(4, 2, 8, 3 * 256 * 256), (32, 2, 8, 3 * 256 * 256), (8, 1, 5, 16 * 16 * 16)]:
        x = torch.randn((B, T, F), device=DEV)
        mo = torch.randn((B * nfe, T, F), device=DEV)
        mi = torch.empty((B * nfe, T, F), device=DEV, dtype=torch.bfloat16)
        nh = torch.randn((B, T, F), device=DEV)
        upd = np.zeros((B * nfe, T), dtype=sp.UPDATE_DTYPE)
        upd["a"], upd["b"], upd["w"], upd["generate"] = 0.9, 0.1, 1.0, 1
        upd["generate"][:, 0] = 0
        prep = np.zeros((B * nfe, T), dtype=sp.PREPARE_DTYPE)
        prep["mode"][0::nfe, 0] = 1
        prep["qa"], prep["qb"] = 0.3, 0.9
        prep["noise_row"] = np.arange(B * nfe)[:, None] // nfe
        ud, pd = sp.to_device_bytes(upd, DEV), sp.to_device_bytes(prep, DEV)
        us = timeit(lambda: ops.sampler_step_hg(x, mo, mi, ud, pd, None, nh, None, B, nfe, T), iters)
        n = B * T * F
        by = n * (nfe * 4 + 4 + 4 + nfe * 2) + B * F * 4        # history-noise read only for the context frame
        print(f"sampler_step_hg B={B} nfe={nfe} T={T} F={F}: {us:7.1f} us  {by / us / 1e3:7.1f} GB/s  ({by / 1e6:.0f} MB)")


def bench_uvit_gemm(iters):
    """GEMM / implicit-GEMM conv shapes of one RE10K-size UViT3DPose forward at batch 4 (8 rows, 64 images)."""
    tot_f, tot_t = 0.0, 0.0
    print("-- convs (n_img, H, W, Cin, Cout, epilogue, count per forward)")
    for (n, H, Cin, Cout, epi, cnt) in [(64, 128, 128, 128, ops.EPI_BF16, 6), (64, 128, 128, 128, ops.EPI_RESID_F32, 6),
                                        (64, 64, 256, 256, ops.EPI_BF16, 6), (64, 64, 256, 256, ops.EPI_RESID_F32, 6),
                                        (64, 64, 128, 256, ops.EPI_F32, 1), (64, 32, 256, 576, ops.EPI_F32, 1),
                                        (64, 16, 576, 1152, ops.EPI_F32, 1), (64, 16, 1152, 576, ops.EPI_F32, 1),
                                        (64, 32, 576, 256, ops.EPI_F32, 1), (64, 64, 256, 128, ops.EPI_F32, 1)]:
        x = torch.randn((n, H, H, Cin), device=DEV).to(torch.bfloat16)
        w = (torch.randn((Cout, 3, 3, Cin), device=DEV) / math.sqrt(9 * Cin)).to(torch.bfloat16)
        bias = torch.randn((Cout,), device=DEV)
        f32 = epi in (ops.EPI_F32, ops.EPI_RESID_F32)
        out = torch.empty((n * H * H, Cout), device=DEV, dtype=torch.float32 if f32 else torch.bfloat16)
        kw = dict(bias=bias)
        if epi == ops.EPI_RESID_F32:
            kw["resid"] = out
        us = timeit(lambda: ops.conv3x3_bf16(x, w, out, epi, **kw), iters)
        fl = 2.0 * n * H * H * Cout * 9 * Cin
        tot_f += fl * cnt
        tot_t += us * cnt
        print(f"conv {H:3d}x{H:<3d} {Cin:4d}->{Cout:<4d} epi={epi}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  x{cnt}")
    print("-- linears (M, N, K, epilogue, count per forward)")
    for (M, N, K, epi, cnt) in [(65536, 1728, 576, ops.EPI_BF16, 12), (65536, 2304, 576, ops.EPI_SILU_BF16, 12),
                                (65536, 576, 576, ops.EPI_RESID_F32, 12), (65536, 576, 2304, ops.EPI_RESID_F32, 12),
                                (65536, 576, 2880, ops.EPI_RESID_F32, 0),
                                (16384, 3456, 1152, ops.EPI_BF16, 20), (16384, 4608, 1152, ops.EPI_SILU_BF16, 20),
                                (16384, 1152, 1152, ops.EPI_RESID_F32, 20), (16384, 1152, 4608, ops.EPI_RESID_F32, 20),
                                (16384, 1152, 5760, ops.EPI_RESID_F32, 0)]:
        a = torch.randn((M, K), device=DEV).to(torch.bfloat16)
        w = (torch.randn((N, K), device=DEV) / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn((N,), device=DEV)
        f32 = epi in (ops.EPI_F32, ops.EPI_RESID_F32)
        out = torch.empty((M, N), device=DEV, dtype=torch.float32 if f32 else torch.bfloat16)
        kw = dict(bias=bias)
        if epi == ops.EPI_RESID_F32:
            kw["resid"] = out
        us = timeit(lambda: ops.gemm_bf16(a, w, out, epi, **kw), iters)
        fl = 2.0 * M * N * K
        tot_f += fl * cnt
        tot_t += us * cnt
        print(f"gemm M={M} N={N:4d} K={K:4d} epi={epi}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  x{cnt}")
    print(f"per forward: {tot_t / 1e3:.2f} ms, {tot_f / tot_t / 1e6:.1f} TFLOP/s average")


def bench_uvit(iters):
    """HBM-bound U-ViT3DPose kernels at the RE10K level-0 / level-2 sizes of a batch-4 (8-row) forward."""
    n, HW, C = 64, 16384, 128
    x32 = torch.randn((n * HW, C), device=DEV)
    x16 = x32.to(torch.bfloat16)
    out = torch.empty((n * HW, C), device=DEV, dtype=torch.bfloat16)
    gamma, beta = torch.randn((C,), device=DEV), torch.randn((C,), device=DEV)
    sums = torch.empty((n, 32, 3), dtype=torch.float64, device=DEV)
    mod_img = torch.randn((n, 4 * C), device=DEV)
    mod_pix = torch.randn((32 * HW, 2 * C), device=DEV).to(torch.bfloat16)
    img_map = torch.tensor([(-1 if i % 16 < 8 else (i // 16) * 8 + i % 8) for i in range(n)], dtype=torch.int32, device=DEV)
    gb = lambda b, us: b / us / 1e3
    us = timeit(lambda: ops.groupnorm_stats(x32, sums, n, HW, C), iters)
    print(f"groupnorm_stats f32  [{n}x{HW}x{C}]: {us:7.1f} us  {gb(n * HW * C * 4, us):7.1f} GB/s")
    us = timeit(lambda: ops.groupnorm_stats(x16, sums, n, HW, C), iters)
    print(f"groupnorm_stats bf16 [{n}x{HW}x{C}]: {us:7.1f} us  {gb(n * HW * C * 2, us):7.1f} GB/s")
    us = timeit(lambda: ops.groupnorm_silu_bf16(x32, sums, gamma, beta, out, n, HW, C), iters)
    print(f"groupnorm_silu f32->bf16:            {us:7.1f} us  {gb(n * HW * C * 6, us):7.1f} GB/s")
    us = timeit(lambda: ops.groupnorm_silu_bf16(x16, sums, gamma, beta, out, n, HW, C, mod_img=mod_img, scale_col=0,
                                                shift_col=C, mod_pix=mod_pix, img_map=img_map), iters)
    print(f"groupnorm_silu bf16+FiLM(half rows): {us:7.1f} us  {gb(n * HW * C * (4 + 2), us):7.1f} GB/s")
    us = timeit(lambda: ops.cast_bf16(x32, out), iters)
    print(f"cast_bf16:                           {us:7.1f} us  {gb(n * HW * C * 6, us):7.1f} GB/s")
    for heads, dh, T, g in [(9, 64, 8, 32), (9, 128, 8, 16)]:
        D, Ntok = heads * dh, T * g * g
        M = 8 * Ntok
        qkv = torch.randn((M, 3 * D), device=DEV).to(torch.bfloat16)
        qw, kw = torch.randn((dh,), device=DEV), torch.randn((dh,), device=DEV)
        table = torch.randn((Ntok, dh // 2, 2), device=DEV)
        us = timeit(lambda: ops.qk_norm_rope(qkv, qw, kw, table, Ntok, heads, dh, 0.18), iters)
        print(f"qk_norm_rope M={M} heads={heads} d={dh}: {us:7.1f} us  {gb(M * 2 * D * 4, us):7.1f} GB/s")
        x = torch.randn((M, D), device=DEV)
        w = torch.randn((D,), device=DEV)
        mi = torch.randn((64, 4 * D), device=DEV)
        mp = torch.randn((32 * g * g, 2 * D), device=DEV).to(torch.bfloat16)
        o = torch.empty((M, D), device=DEV, dtype=torch.bfloat16)
        us = timeit(lambda: ops.rmsnorm_film_bf16(x, w, mi, 0, D, g * g, o, mod_pix=mp, img_map=img_map), iters)
        print(f"rmsnorm_film M={M} D={D}:           {us:7.1f} us  {gb(M * D * (4 + 2 + 2), us):7.1f} GB/s")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="?", default="all")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    if a.which in ("attn", "all"):
        bench_attn(a.iters)
    if a.which in ("gemm", "all"):
        bench_gemm(a.iters)
    if a.which in ("norm", "all"):
        bench_norm(a.iters)
    if a.which in ("matrix", "all"):
        bench_matrix(a.iters)
    if a.which in ("gemm_ditb",):
        bench_gemm_ditb(a.iters)
    if a.which in ("sampler", "all"):
        bench_sampler(a.iters)
    if a.which in ("uvit", "all"):
        bench_uvit(a.iters)
    if a.which in ("uvit_gemm", "all"):
        bench_uvit_gemm(a.iters)
