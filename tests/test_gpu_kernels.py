"""-m gpu: every CUDA kernel, called through the C ABI, against a plain PyTorch fp32 reference of the same op
(inputs rounded to bf16 where the kernel consumes bf16, so only accumulation order / output rounding differ)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from dfot_b200 import ops  # noqa: E402
from dfot_b200.algorithms.dfot import sampling_plan as sp  # noqa: E402
from dfot_b200.algorithms.dfot.backbones.dit.dit3d import rope_cos_sin_table  # noqa: E402
import k4_emulation  # noqa: E402

DEV = "cuda"


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def rel_err(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp(min=1e-12)).item()


# ---------------------------------------------------------------- K2 GEMM
GEMM_SHAPES = [(128, 256, 64), (40, 64, 256), (1000, 192, 64), (333, 1152, 1152), (2560, 3456, 1152),
               (512, 16, 1152), (40, 4608, 64), (130, 72, 8), (2048, 1152, 4608), (256, 136, 200)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bias_f32_and_bf16(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn((M, K), generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g) / math.sqrt(K)).to(DEV).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g).to(DEV)
    ref = a.float() @ w.float().t() + bias
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm_bf16(a, w, out, ops.EPI_F32, bias=bias)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item()), (out - ref).abs().max()
    out16 = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    ops.gemm_bf16(a, w, out16, ops.EPI_BF16, bias=bias)
    assert rel_err(out16, ref) < 5e-3


@pytest.mark.parametrize("P,M", [(64, 640), (16, 640), (40, 640), (4, 640), (20, 640), (8, 600), (16, 20480), (48, 18432)])
def test_gemm_activations_and_gate_residual(P, M):
    """(incl. frames smaller than a 32-row epilogue chunk — 16 tokens per frame are the DMLab / Minecraft DiT shapes —
    ragged M, and sizes that take the CTA-pair kernel)"""
    N, K = 1152, 256
    g = torch.Generator().manual_seed(5)
    a = torch.randn((M, K), generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g) / math.sqrt(K)).to(DEV).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g).to(DEV)
    pre = a.float() @ w.float().t() + bias
    o = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    ops.gemm_bf16(a, w, o, ops.EPI_GELU_BF16, bias=bias)
    assert (o.float() - torch.nn.functional.gelu(pre, approximate="tanh")).abs().max().item() < 2e-2
    ops.gemm_bf16(a, w, o, ops.EPI_SILU_BF16, bias=bias)
    assert (o.float() - torch.nn.functional.silu(pre)).abs().max().item() < 2e-2
    frames = M // P
    mod = torch.randn((frames, 3 * N), generator=g).to(DEV)
    resid = torch.randn((M, N), generator=g).to(DEV)
    of = torch.empty((M, N), device=DEV)
    ops.gemm_bf16(a, w, of, ops.EPI_GATE_RESID_F32, bias=bias, resid=resid, gate=mod[:, 2 * N:], ld_gate=3 * N,
                  tokens_per_frame=P)
    ref = resid + mod[:, 2 * N:].repeat_interleave(P, 0) * pre
    assert (of - ref).abs().max().item() < 5e-3


@pytest.mark.parametrize("heads,dh,T,gh,gw", [(4, 64, 8, 8, 8), (16, 72, 5, 16, 16), (2, 128, 3, 4, 4)])
def test_gemm_qkv_rope_epilogue(heads, dh, T, gh, gw):
    D, R = heads * dh, 2
    Ntok = T * gh * gw
    M = R * Ntok
    g = torch.Generator().manual_seed(dh)
    a = torch.randn((M, D), generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn((3 * D, D), generator=g) / math.sqrt(D)).to(DEV).to(torch.bfloat16)
    bias = torch.randn((3 * D,), generator=g).to(DEV)
    cs = rope_cos_sin_table(dh, (T, gh, gw)).to(DEV)
    qs = 0.37
    out = torch.empty((M, 3 * D), device=DEV, dtype=torch.bfloat16)
    ops.gemm_bf16(a, w, out, ops.EPI_QKV_ROPE_BF16, bias=bias, rope_cs=cs, tokens_per_sample=Ntok, model_dim=D,
                  head_dim=dh, q_scale=qs)
    pre = (a.float() @ w.float().t() + bias).reshape(R, Ntok, 3, heads, dh)
    cos = cs[..., 0].repeat_interleave(2, -1)[None, :, None, None, :]
    sin = cs[..., 1].repeat_interleave(2, -1)[None, :, None, None, :]
    pairs = pre.reshape(*pre.shape[:-1], -1, 2)
    rot = torch.stack((-pairs[..., 1], pairs[..., 0]), -1).reshape(pre.shape)
    roped = pre * cos + rot * sin
    ref = pre.clone()
    ref[:, :, 0] = roped[:, :, 0] * qs
    ref[:, :, 1] = roped[:, :, 1]
    assert (out.float().reshape(ref.shape) - ref).abs().max().item() < 3e-2
    assert rel_err(out.reshape(ref.shape), ref) < 5e-3


# ---------------------------------------------------------------- K3 attention
@pytest.mark.parametrize("R,heads,dh,N", [(2, 4, 64, 512), (1, 1, 64, 64), (2, 16, 72, 1280), (1, 12, 64, 576),
                                          (1, 9, 128, 2048), (3, 2, 72, 200), (1, 2, 64, 1), (2, 3, 72, 128),
                                          (3, 2, 128, 129), (1, 2, 64, 256), (5, 7, 72, 384), (1, 2, 64, 8192),
                                          (8, 16, 72, 1280),
                                          # more pair items than SMs with an even tile count: the last wave is handed
                                          # out as single query tiles (split tail) — dual-issuer, single-issuer, d = 128
                                          (4, 12, 64, 1024), (3, 9, 128, 2048), (5, 16, 72, 512),
                                          # head_dim 72 runs 112-key KV tiles: whole tiles, one key over, T = 17 stress
                                          (2, 2, 72, 224), (1, 2, 72, 225), (2, 3, 72, 337), (1, 4, 72, 4352)])
def test_attention_matches_sdpa(R, heads, dh, N):
    D = heads * dh
    g = torch.Generator().manual_seed(N + dh)
    qkv = torch.randn((R * N, 3 * D), generator=g).to(DEV).to(torch.bfloat16)
    out = torch.full((R * N, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, R, N, heads, dh)
    q, k, v = qkv.float().reshape(R, N, 3, heads, dh).permute(2, 0, 3, 1, 4).unbind(0)
    # the kernel expects q pre-multiplied by scale*log2(e) and uses exp2: softmax2(q k^T) == softmax(ln2 * q k^T)
    w = torch.softmax(q @ k.transpose(-1, -2) * math.log(2.0), dim=-1)
    ref = (w @ v).transpose(1, 2).reshape(R * N, D)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() < 3e-2, (out.float() - ref).abs().max()
    assert rel_err(out, ref) < 1e-2


@pytest.mark.parametrize("R,heads,dh,N", [(2, 9, 64, 1024), (1, 9, 128, 512), (2, 3, 64, 200), (1, 2, 128, 129),
                                          (1, 2, 64, 100), (3, 4, 72, 384), (4, 12, 64, 1024), (3, 9, 128, 2048), (2, 3, 72, 1233)])
def test_attention_bounded_scores_matches_sdpa(R, heads, dh, N):
    """QK-normalised operands with a declared score bound: the kernel may drop the running maximum (p = 2^s); the
    result must still be the softmax attention.  Scores reach the declared bound on the diagonal (q == k direction)."""
    D = heads * dh
    g = torch.Generator().manual_seed(N + dh + 1)
    t = torch.randn((R, N, 3, heads, dh), generator=g)
    wq, wk = 1.7, 1.3                                               # norm weights
    scale = math.log2(math.e) / math.sqrt(dh)
    t[:, :, 1] = t[:, :, 1] / t[:, :, 1].pow(2).mean(-1, keepdim=True).sqrt() * wk
    t[:, :, 0] = t[:, :, 1] / wk * wq * scale                       # q parallel to its own key: score = the bound
    t[:, : N // 2, 0] = torch.randn((R, N // 2, heads, dh), generator=g) * scale   # ... for half of the queries
    bound = 1.05 * math.sqrt(dh) * wq * wk * math.log2(math.e)
    qkv = t.reshape(R * N, 3 * D).to(DEV).to(torch.bfloat16)
    q, k, v = qkv.float().reshape(R, N, 3, heads, dh).permute(2, 0, 3, 1, 4).unbind(0)
    s = q @ k.transpose(-1, -2)
    assert s.abs().max().item() <= bound
    ref = (torch.softmax(s * math.log(2.0), dim=-1) @ v).transpose(1, 2).reshape(R * N, D)
    out = torch.full((R * N, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, R, N, heads, dh, score_bound=bound)
    assert torch.isfinite(out.float()).all()
    # (absolute 3e-2 for |ref| <= 1, relative above: one bf16 ulp of an output in [4, 8) is already 3.1e-2)
    assert ((out.float() - ref).abs() <= 3e-2 * ref.abs().clamp(min=1.0)).all() and rel_err(out, ref) < 1e-2


@pytest.mark.parametrize("heads,dh,N", [(3, 64, 2048), (2, 128, 1024), (2, 72, 640)])
def test_attention_rescale_stress_is_deterministic(heads, dh, N):
    """Large, growing logits force the lazy-rescale path (O tile rescaled in tensor memory) on most KV tiles; repeated
    launches must agree bit-for-bit (a race between the PV MMA and the P / O updates would show up as jitter)."""
    R, D = 2, heads * dh
    g = torch.Generator().manual_seed(dh)
    qkv = torch.randn((R, N, 3, heads, dh), generator=g)
    qkv[:, :, 1] *= torch.linspace(0.2, 3.0, N)[None, :, None, None]      # later keys carry larger scores
    qkv = qkv.reshape(R * N, 3 * D).to(DEV).to(torch.bfloat16)
    q, k, v = qkv.float().reshape(R, N, 3, heads, dh).permute(2, 0, 3, 1, 4).unbind(0)
    w = torch.softmax(q @ k.transpose(-1, -2) * math.log(2.0), dim=-1)
    ref = (w @ v).transpose(1, 2).reshape(R * N, D)
    outs = []
    for _ in range(6):
        out = torch.full((R * N, D), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.attention(qkv, out, R, N, heads, dh)
        outs.append(out)
    torch.cuda.synchronize()
    assert (outs[0].float() - ref).abs().max().item() < 4e-2
    assert rel_err(outs[0], ref) < 1.5e-2
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


# ---------------------------------------------------------------- K1 adaLN + LayerNorm
@pytest.mark.parametrize("latency_mode", [False, True])    # on: one block per row up to 2048 rows (and D <= 2048)
@pytest.mark.parametrize("M,D,P", [(512, 256, 64), (1280, 1152, 256), (100, 64, 16), (96, 768, 16), (33, 2048, 11),
                                   (4096, 1152, 256), (2049, 2048, 3), (300, 4096, 100)])
def test_adaln_layernorm(M, D, P, latency_mode):
    g = torch.Generator().manual_seed(M + D)
    x = (torch.randn((M, D), generator=g) * 3 + 0.5).to(DEV)
    frames = (M + P - 1) // P
    mod = torch.randn((frames, 6 * D), generator=g).to(DEV)
    y32 = torch.empty((M, D), device=DEV)
    y16 = torch.empty((M, D), device=DEV, dtype=torch.bfloat16)
    ops.set_latency_mode(latency_mode)
    try:
        ops.adaln_layernorm(x, mod, 3 * D, 4 * D, P, y_f32=y32, y_bf16=y16)
    finally:
        ops.set_latency_mode(False)
    f = torch.arange(M, device=DEV) // P
    ref = torch.nn.functional.layer_norm(x, (D,), eps=1e-6) * (1 + mod[f, 4 * D:5 * D]) + mod[f, 3 * D:4 * D]
    assert (y32 - ref).abs().max().item() < 1e-4
    assert (y16.float() - ref).abs().max().item() < 5e-2 and rel_err(y16, ref) < 4e-3


# ---------------------------------------------------------------- K4 fused sampler step vs contract emulation
@pytest.mark.parametrize("out_dt,in_dt", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                          (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("B,nfe,T,shape", [(2, 2, 8, (4, 16, 16)), (8, 1, 5, (16, 16, 16)), (1, 6, 4, (4, 8, 8)),
                                           (1, 2, 8, (3, 64, 64)), (2, 3, 4, (4, 8, 8)), (1, 4, 3, (4, 16, 16))])
def test_sampler_step_hg_matches_contract(out_dt, in_dt, B, nfe, T, shape):
    g = torch.Generator().manual_seed(B * 100 + nfe * 10 + T)
    rng = np.random.default_rng(B + nfe + T)
    R = B * nfe
    x = torch.randn((B, T, *shape), generator=g)
    mo = torch.randn((R, T, *shape), generator=g).to(out_dt)
    nd, nh, ne = (torch.randn((R, T, *shape), generator=g) for _ in range(3))
    upd = np.zeros((R, T), dtype=sp.UPDATE_DTYPE)
    upd["a"], upd["b"] = rng.normal(size=(R, T)), rng.normal(size=(R, T))
    upd["sigma"] = np.where(rng.random((R, T)) < 0.5, 0.0, rng.normal(size=(R, T)))
    upd["w"] = np.where(rng.random((R, T)) < 0.2, 0.0, rng.normal(size=(R, T)))
    upd["clip"] = np.where(rng.random((R, T)) < 0.5, 0.0, 0.7)
    gen = (rng.random((B, T)) < 0.6).astype(np.int32)
    upd["generate"] = np.repeat(gen[:, None], nfe, 1).reshape(R, T)
    prep = np.zeros((R, T), dtype=sp.PREPARE_DTYPE)
    prep["mode"] = rng.integers(0, 3, size=(R, T))
    prep["noise_row"] = rng.integers(0, R, size=(R, T))
    prep["qa"], prep["qb"] = rng.random((R, T)), rng.random((R, T))
    x_ref, mi_ref = x.clone(), torch.empty((R, T, *shape), dtype=in_dt)
    k4_emulation.emulate(x_ref, mo, mi_ref, upd, prep, nd, nh, ne, B, nfe, T)
    xd, mi = x.to(DEV), torch.full((R, T, *shape), float("nan"), device=DEV, dtype=in_dt)
    ops.sampler_step_hg(xd, mo.to(DEV), mi, sp.to_device_bytes(upd, DEV), sp.to_device_bytes(prep, DEV), nd.to(DEV),
                        nh.to(DEV), ne.to(DEV), B, nfe, T)
    assert (xd.cpu() - x_ref).abs().max().item() < 1e-4
    tol = 1e-5 if in_dt == torch.float32 else 4e-2
    assert (mi.float().cpu() - mi_ref.float()).abs().max().item() < tol
    # prepare-only and update-only launches
    x2 = x.to(DEV)
    ops.sampler_step_hg(x2, None, mi, None, sp.to_device_bytes(prep, DEV), None, nh.to(DEV), ne.to(DEV), B, nfe, T)
    assert torch.equal(x2.cpu(), x)
    x3 = x.to(DEV)
    ops.sampler_step_hg(x3, mo.to(DEV), None, sp.to_device_bytes(upd, DEV), None, nd.to(DEV), None, None, B, nfe, T)
    assert (x3.cpu() - x_ref).abs().max().item() < 1e-4


# ---------------------------------------------------------------- glue kernels
def test_noise_features_patchify_unpatchify_silu():
    g = torch.Generator().manual_seed(0)
    lv = torch.randint(0, 1000, (40,), generator=g)
    out = torch.empty((40, 256), device=DEV, dtype=torch.bfloat16)
    ops.noise_features(lv.to(DEV), out)
    fr = torch.exp(-math.log(10000) * torch.arange(128, dtype=torch.float32) / 128)
    ang = lv[:, None].float() * fr
    ref = torch.cat([ang.cos(), ang.sin()], -1)
    assert (out.float().cpu() - ref).abs().max().item() < 1e-2
    lvf = torch.randn((40,), generator=g) * 2
    freqs, phases = torch.randn((256,), generator=g) * 6, torch.rand((256,), generator=g) * 6
    ops.noise_features(lvf.to(DEV), out, freqs.to(DEV), phases.to(DEV))
    ref = torch.cos(lvf[:, None] * freqs + phases) * math.sqrt(2)
    assert (out.float().cpu() - ref).abs().max().item() < 1e-2
    # patchify / unpatchify
    for C, H, W, p in [(4, 16, 16, 2), (16, 16, 16, 1), (32, 8, 8, 2)]:
        fr_ = 6
        x = torch.randn((fr_, C, H, W), generator=g)
        kk = C * p * p
        pat = torch.zeros((fr_ * (H // p) * (W // p), (kk + 7) // 8 * 8), device=DEV, dtype=torch.bfloat16)
        if pat.shape[1] == kk:
            ops.patchify_bf16(x.to(DEV), pat, fr_, C, H, W, p)
            ref = torch.nn.functional.unfold(x, p, stride=p).transpose(1, 2).reshape(-1, kk)
            assert torch.equal(pat.float().cpu(), ref.to(torch.bfloat16).float())
        tok = torch.randn((fr_ * (H // p) * (W // p), (kk + 7) // 8 * 8), generator=g)
        xo = torch.empty((fr_, C, H, W), device=DEV)
        ops.unpatchify(tok.to(DEV), xo, fr_, C, H, W, p)
        ref = tok[:, :kk].reshape(fr_, H // p, W // p, p, p, C).permute(0, 5, 1, 3, 2, 4).reshape(fr_, C, H, W)
        assert torch.equal(xo.cpu(), ref)
    a, b = torch.randn((24, 64), generator=g), torch.randn((24, 64), generator=g)
    mask = torch.tensor([1, 0, 1], dtype=torch.uint8)
    o = torch.empty((24, 64), device=DEV, dtype=torch.bfloat16)
    ops.silu_sum_bf16(a.to(DEV), b.to(DEV), mask.to(DEV), 8, o)
    keep = (1 - mask.float()).repeat_interleave(8)[:, None]
    assert (o.float().cpu() - torch.nn.functional.silu(a + keep * b)).abs().max().item() < 2e-2


# ---------------------------------------------------------------- matrix-attention patch kernels (the u factors)
@pytest.mark.parametrize("R,L,P,Mc,D", [(2, 4, 16, 1, 64), (1, 5, 64, 2, 512), (3, 8, 256, 1, 768), (1, 3, 9, 5, 132),
                                        (2, 16, 16, 16, 64)])
def test_patch_mix_and_expand(R, L, P, Mc, D):
    g = torch.Generator().manual_seed(R * 131 + P)
    y = torch.randn((R * L * P, D), generator=g)
    u = torch.randn((P, Mc), generator=g) / math.sqrt(P)
    out = torch.empty((R * Mc * L, D), device=DEV, dtype=torch.bfloat16)
    ops.patch_mix_bf16(y.to(DEV), u.to(DEV), out, R, L, P, Mc)
    ref = torch.einsum("nc,rlnd->rcld", u.double(), y.reshape(R, L, P, D).double()).reshape(R * Mc * L, D).float()
    assert (out.float().cpu() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item() + 1e-6   # bf16 rounding
    out2 = torch.empty_like(out)
    ops.patch_mix_bf16(y.to(DEV), u.to(DEV), out2, R, L, P, Mc)
    assert torch.equal(out, out2)                                                                  # fixed-order reduction
    # the way back: x = y + gate[frame] * (pu^T z + pb)
    z = torch.randn((R * Mc * L, D), generator=g)
    pu = torch.randn((Mc, P), generator=g)
    pb = torch.randn((P, D), generator=g)
    mod = torch.randn((R * L, 3 * D + 8), generator=g)                 # the gate is a column slab of a wider matrix
    for bias in (pb, None):
        x = torch.empty((R * L * P, D), device=DEV)
        ops.patch_expand_gate_resid(x, y.to(DEV), z.to(DEV), pu.to(DEV), None if bias is None else bias.to(DEV),
                                    mod.to(DEV)[:, 2 * D:], mod.shape[1], R, L, P, Mc)
        s = torch.einsum("cn,rcld->rlnd", pu.double(), z.reshape(R, Mc, L, D).double())
        if bias is not None:
            s = s + bias.double()
        ref = y.reshape(R, L, P, D).double() + mod[:, 2 * D:3 * D].reshape(R, L, 1, D).double() * s
        assert (x.cpu().double() - ref.reshape(-1, D)).abs().max().item() < 1e-4
    # bare form (a MatrixCrossDiTBlock's attn1 output): no residual, gate 1
    x = torch.full((R * L * P, D), float("nan"), device=DEV)
    ops.patch_expand_gate_resid(x, None, z.to(DEV), pu.to(DEV), pb.to(DEV), None, 0, R, L, P, Mc)
    ref = torch.einsum("cn,rcld->rlnd", pu.double(), z.reshape(R, Mc, L, D).double()) + pb.double()
    assert (x.cpu().double() - ref.reshape(-1, D)).abs().max().item() < 1e-4


# ---------------------------------------------------------------- split-K GEMM + its fused consumer (latency regime)
@pytest.mark.parametrize("M,N,K,S", [(256, 768, 3072, 6), (256, 768, 768, 3), (100, 64, 128, 2), (512, 1152, 4608, 2),
                                     (128, 64, 64, 1), (300, 256, 1024, 8), (256, 1536, 520, 5)])
def test_gemm_splitk_partials_and_fused_consumer(M, N, K, S):
    g = torch.Generator().manual_seed(M + N + K)
    a = bf16_round(torch.randn((M, K), generator=g))
    w = bf16_round(torch.randn((N, K), generator=g) / math.sqrt(K))
    parts = torch.full((M, S * N), float("nan"), device=DEV)
    ops.gemm_bf16_splitk(a.to(DEV).to(torch.bfloat16), w.to(DEV).to(torch.bfloat16), parts, S)
    kb = -(-K // 64)
    p = parts.cpu().view(M, S, N)
    for s in range(S):                                              # every split holds exactly its share of the k-blocks
        k0, k1 = 64 * (s * kb // S), min(K, 64 * ((s + 1) * kb // S))
        ref = a[:, k0:k1].double() @ w[:, k0:k1].double().t()
        assert (p[:, s].double() - ref).abs().max().item() < 2e-3, s
    full = a.double() @ w.double().t()
    assert (p.double().sum(1) - full).abs().max().item() < 2e-3
    # consumer: x = resid + gate * (sum + bias); y = LN(x) * (1 + scale) + shift
    tpf = 16 if M % 16 == 0 else 1
    frames = -(-M // tpf)
    resid, bias = torch.randn((M, N), generator=g), torch.randn((N,), generator=g)
    mod = torch.randn((frames, 3 * N + 8), generator=g) * 0.5
    x = torch.empty((M, N), device=DEV)
    y32 = torch.empty((M, N), device=DEV)
    y16 = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    ops.splitk_gate_resid_adaln(parts, S, bias.to(DEV), resid.to(DEV), mod.to(DEV), 2 * N, 0, N, tpf, x_out=x, y_f32=y32,
                                y_bf16=y16)
    fr = torch.arange(M) // tpf
    xr = resid.double() + mod[fr, 2 * N:3 * N].double() * (p.double().sum(1) + bias.double())
    yr = torch.nn.functional.layer_norm(xr, (N,), eps=1e-6) * (1 + mod[fr, N:2 * N].double()) + mod[fr, :N].double()
    assert (x.cpu().double() - xr).abs().max().item() < 1e-4
    assert (y32.cpu().double() - yr).abs().max().item() < 1e-3
    assert (y16.float().cpu().double() - yr).abs().max().item() < 2e-2 * max(1.0, yr.abs().max().item())
    x2 = torch.empty((M, N), device=DEV)                            # no norm: x only, no bias
    ops.splitk_gate_resid_adaln(parts, S, None, resid.to(DEV), mod.to(DEV), 2 * N, -1, -1, tpf, x_out=x2)
    xr2 = resid.double() + mod[fr, 2 * N:3 * N].double() * p.double().sum(1)
    assert (x2.cpu().double() - xr2).abs().max().item() < 1e-4


# ---------------------------------------------------------------- residual base rebuilt in the gate epilogue
@pytest.mark.parametrize("M,N,K,tpf", [(512, 768, 768, 256), (300, 1152, 320, 40), (256, 768, 3072, 16), (130, 64, 64, 1),
                                       (10240, 1152, 1152, 256), (384, 256, 256, 32)])
def test_gate_lnresid_epilogue_equals_the_stored_residual_base(M, N, K, tpf):
    """DFOT_EPI_GATE_LNRESID_F32 (K1 writes bf16 + row statistics, the GEMM epilogue rebuilds modulate(LN(x)) from x in place)
    against the two-buffer path (K1 stores the fp32 copy, DFOT_EPI_GATE_RESID_F32 reads it): bit-identical."""
    g = torch.Generator().manual_seed(M + N + K + tpf)
    frames = -(-M // tpf)
    x = torch.randn((M, N), generator=g).to(DEV) * 2 + 0.3
    mod = (torch.randn((frames, 3 * N + 8), generator=g) * 0.5).to(DEV)
    a = torch.randn((M, K), generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g) / math.sqrt(K)).to(DEV).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g).to(DEV)
    # stored copy
    y32, y16 = torch.empty_like(x), torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    ops.adaln_layernorm(x, mod, 0, N, tpf, y_f32=y32, y_bf16=y16)
    want = torch.empty_like(x)
    ops.gemm_bf16(a, w, want, ops.EPI_GATE_RESID_F32, bias=bias, resid=y32, gate=mod[:, 2 * N:], ld_gate=mod.shape[1],
                  tokens_per_frame=tpf)
    # rebuilt
    stats, y16b = torch.empty((M, 2), device=DEV), torch.empty_like(y16)
    ops.adaln_layernorm(x, mod, 0, N, tpf, y_bf16=y16b, stats=stats)
    assert torch.equal(y16, y16b)
    ref_mean = x.double().mean(1)
    assert (stats[:, 0].double() - ref_mean).abs().max().item() < 1e-5
    assert (stats[:, 1].double() - 1 / torch.sqrt(x.double().var(1, unbiased=False) + 1e-6)).abs().max().item() < 1e-4
    xs = x.clone()
    ops.gemm_bf16(a, w, xs, ops.EPI_GATE_LNRESID_F32, bias=bias, resid=xs, gate=mod[:, 2 * N:], ld_gate=mod.shape[1],
                  tokens_per_frame=tpf, ln_stats=stats, ln_shift=mod[:, 0:], ln_scale=mod[:, N:])
    assert torch.equal(xs, want), (xs - want).abs().max().item()
