"""-m gpu: the U-ViT3DPose kernels (implicit-GEMM 3x3 conv, GroupNorm/RMSNorm + FiLM, q/k norm + RoPE, pooling,
upsampling, ray encoding), called through the C ABI, against plain PyTorch fp32 references of the same ops."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from dfot_b200 import ops  # noqa: E402
from dfot_b200.algorithms.dfot.backbones.dit.dit3d import rope_cos_sin_table  # noqa: E402

DEV = "cuda"


def bf16_close(out, ref, atol=2e-3):
    """bf16 output vs fp32 reference: one bf16 ulp (2^-8 relative) plus a small absolute slack."""
    return bool(((out.float() - ref).abs() <= ref.abs() * 2.0 ** -7 + atol).all())


def rel_err(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp(min=1e-12)).item()


# (n_img, H, W, Cin, Cout): RE10K levels (shrunk batch), the tiny golden config, odd channel counts
CONV_SHAPES = [(2, 128, 128, 128, 128), (3, 64, 64, 256, 256), (2, 64, 64, 128, 256), (4, 32, 32, 256, 576),
               (8, 16, 16, 576, 1152), (8, 16, 16, 1152, 576), (4, 16, 16, 32, 32), (4, 8, 8, 32, 32),
               (4, 4, 4, 32, 64), (8, 2, 2, 64, 128), (5, 2, 2, 128, 64), (3, 8, 8, 64, 32), (1, 256, 256, 8, 16)]


@pytest.mark.parametrize("n,H,W,Cin,Cout", CONV_SHAPES)
def test_conv3x3_implicit_gemm(n, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(n * 1000 + H + Cin + Cout)
    x = torch.randn((n, Cin, H, W), generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn((Cout, Cin, 3, 3), generator=g) / math.sqrt(9 * Cin)).to(DEV).to(torch.bfloat16)
    bias = torch.randn((Cout,), generator=g).to(DEV)
    ref = F.conv2d(x.float(), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(n * H * W, Cout)
    x_cl = x.permute(0, 2, 3, 1).contiguous()
    w_cl = w.permute(0, 2, 3, 1).contiguous()
    out = torch.full((n * H * W, Cout), float("nan"), device=DEV)
    ops.conv3x3_bf16(x_cl, w_cl, out, ops.EPI_F32, bias=bias)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() <= 3e-3 * max(1.0, ref.abs().max().item()), (out - ref).abs().max()
    resid = torch.randn((n * H * W, Cout), generator=g).to(DEV)
    out2 = torch.empty_like(out)
    ops.conv3x3_bf16(x_cl, w_cl, out2, ops.EPI_RESID_F32, bias=bias, resid=resid)
    assert (out2 - (ref + resid)).abs().max().item() <= 3e-3 * max(1.0, ref.abs().max().item())
    out16 = torch.empty((n * H * W, Cout), device=DEV, dtype=torch.bfloat16)
    ops.conv3x3_bf16(x_cl, w_cl, out16, ops.EPI_BF16, bias=bias)
    assert rel_err(out16, ref) < 5e-3


@pytest.mark.parametrize("B,T,H,W,Cin,Cout,kt", [(2, 5, 16, 16, 64, 128, 3), (1, 9, 32, 32, 128, 64, 3), (2, 3, 8, 8, 64, 64, 3),
                                                 (1, 3, 4, 4, 128, 128, 3), (1, 17, 64, 64, 64, 8, 3), (2, 4, 32, 32, 64, 64, 1),
                                                 (3, 5, 128, 128, 32, 32, 3)])
def test_conv3d_causal_implicit_gemm(B, T, H, W, Cin, Cout, kt):
    """Causal kt x 3 x 3 convolution (the reference VideoVAE's PaddedConv3D, conv.py:98-108) on the implicit-GEMM
    kernel: clips laid out with kt-1 leading copies of their first frame; the kt-1 output slots that straddle two clips
    are junk by construction and skipped."""
    g = torch.Generator().manual_seed(B * 100 + T * 10 + Cin + Cout)
    x = torch.randn((B, Cin, T, H, W), generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn((Cout, Cin, kt, 3, 3), generator=g) / math.sqrt(9 * kt * Cin)).to(DEV).to(torch.bfloat16)
    bias = torch.randn((Cout,), generator=g).to(DEV)
    xp = torch.cat([x[:, :, :1].repeat(1, 1, kt - 1, 1, 1), x], 2).float()                  # conv.py:100-104
    ref = F.conv3d(xp, w.float(), bias, padding=(0, 1, 1)).permute(0, 2, 3, 4, 1)             # [B, T, H, W, Cout]
    P = kt - 1
    frames = torch.cat([x[:, :, :1].repeat(1, 1, P, 1, 1), x], 2).permute(0, 2, 3, 4, 1)      # [B, T+P, H, W, Cin]
    x_cl = frames.reshape(B * (T + P), H, W, Cin).contiguous()
    w_cl = w.permute(0, 2, 3, 4, 1).contiguous()
    n_out = B * (T + P) - P
    out = torch.full((n_out * H * W, Cout), float("nan"), device=DEV)
    ops.conv3d_causal_bf16(x_cl, w_cl, out, ops.EPI_F32, bias=bias)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    tol = 3e-3 * max(1.0, ref.abs().max().item())
    got = out.reshape(n_out, H, W, Cout)
    for b in range(B):
        sl = got[b * (T + P): b * (T + P) + T]
        assert (sl - ref[b]).abs().max().item() <= tol, (b, (sl - ref[b]).abs().max())
    resid = torch.randn((n_out * H * W, Cout), generator=g).to(DEV)
    out2 = torch.empty_like(out)
    ops.conv3d_causal_bf16(x_cl, w_cl, out2, ops.EPI_RESID_F32, bias=bias, resid=resid)
    assert (out2 - out - resid).abs().max().item() <= 1e-5 * max(1.0, out.abs().max().item())


def fixed_point_sums(ws, k):
    """(sum, sum of squares) of a GroupNorm workspace: 64-bit fixed point in 2^-32 units (csrc/uvit.cu)."""
    return ws.reshape(-1)[:k].view(torch.int64).double() / 2.0 ** 32


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_groupnorm_statistics_are_deterministic_and_batch_invariant(dt):
    """Fixed-point accumulation: the statistics of an image do not depend on the run, on how many images share the launch
    (which changes the pixel split over blocks) or on the path that produced them twice (stand-alone pass)."""
    n, HW, C = 6, 4096, 128
    g = torch.Generator().manual_seed(5)
    x = (torch.randn((n, HW, C), generator=g) * 3 + 0.3).to(DEV).to(dt)
    full = torch.empty((n, 32, 3), dtype=torch.float64, device=DEV)
    ops.groupnorm_stats(x, full, n, HW, C)
    again = torch.empty_like(full)
    for _ in range(3):
        ops.groupnorm_stats(x, again, n, HW, C)
        assert torch.equal(full.view(torch.int64), again.view(torch.int64))
    for i in (0, 3, 5):                       # one image alone: other grid, other slab size — same bits
        one = torch.empty((1, 32, 3), dtype=torch.float64, device=DEV)
        ops.groupnorm_stats(x[i:i + 1].contiguous(), one, 1, HW, C)
        assert torch.equal(fixed_point_sums(one, 64), fixed_point_sums(full, n * 64)[i * 64:(i + 1) * 64])


def test_conv_groupnorm_side_output_is_batch_invariant():
    """The statistics riding on the conv epilogue: image i inside a batch of 5 == image i convolved alone, bit for bit,
    and so is the convolution output itself (what makes a forward independent of the rows it is batched with)."""
    n, H, C = 5, 32, 128
    g = torch.Generator().manual_seed(9)
    x = torch.randn((n, H, H, C), generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn((C, 3, 3, C), generator=g) / math.sqrt(9 * C)).to(DEV).to(torch.bfloat16)
    bias = torch.randn((C,), generator=g).to(DEV)
    side = torch.empty((n, 32, 3), dtype=torch.float64, device=DEV)
    out = torch.empty((n * H * H, C), dtype=torch.bfloat16, device=DEV)
    ops.conv3x3_bf16(x, w, out, ops.EPI_BF16, bias=bias, gn_sums=side)
    for i in (0, 2, 4):
        s1 = torch.empty((1, 32, 3), dtype=torch.float64, device=DEV)
        o1 = torch.empty((H * H, C), dtype=torch.bfloat16, device=DEV)
        ops.conv3x3_bf16(x[i:i + 1].contiguous(), w, o1, ops.EPI_BF16, bias=bias, gn_sums=s1)
        assert torch.equal(o1, out[i * H * H:(i + 1) * H * H])
        assert torch.equal(fixed_point_sums(s1, 64), fixed_point_sums(side, n * 64)[i * 64:(i + 1) * 64])


@pytest.mark.parametrize("n,H,C,epi", [(4, 64, 128, "bf16"), (4, 64, 128, "resid"), (3, 32, 256, "bf16"), (3, 32, 256, "f32"),
                                       (4, 16, 32, "bf16"), (4, 16, 32, "resid"), (6, 8, 64, "bf16"), (2, 16, 1024, "f32")])
def test_conv_groupnorm_side_output(n, H, C, epi):
    """GroupNorm statistics produced by the conv epilogue == a stand-alone statistics pass over the stored output."""
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn((n, H, H, C), generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn((C, 3, 3, C), generator=g) / math.sqrt(9 * C)).to(DEV).to(torch.bfloat16)
    bias = torch.randn((C,), generator=g).to(DEV)
    side = torch.full((n, 32, 3), float("nan"), dtype=torch.float64, device=DEV)
    alone = torch.empty_like(side)
    if epi == "bf16":
        out = torch.empty((n * H * H, C), dtype=torch.bfloat16, device=DEV)
        ops.conv3x3_bf16(x, w, out, ops.EPI_BF16, bias=bias, gn_sums=side)
    else:
        out = torch.empty((n * H * H, C), device=DEV)
        resid = torch.randn((n * H * H, C), generator=g).to(DEV) if epi == "resid" else None
        ops.conv3x3_bf16(x, w, out, ops.EPI_RESID_F32 if epi == "resid" else ops.EPI_F32, bias=bias, resid=resid,
                         gn_sums=side)
    ops.groupnorm_stats(out, alone, n, H * H, C)
    k = n * 32 * 2
    a, b = fixed_point_sums(side, k), fixed_point_sums(alone, k)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-2), (a - b).abs().max()
    fa = side.reshape(-1)[k:].view(torch.float32).reshape(n, 32, 2)
    fb = alone.reshape(-1)[k:].view(torch.float32).reshape(n, 32, 2)
    assert torch.allclose(fa, fb, rtol=1e-3, atol=1e-4)


def test_gemm_resid_epilogue():
    M, N, K = 640, 576, 2880
    g = torch.Generator().manual_seed(11)
    a = torch.randn((M, K), generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g) / math.sqrt(K)).to(DEV).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g).to(DEV)
    resid = torch.randn((M, N), generator=g).to(DEV)
    out = torch.empty((M, N), device=DEV)
    ops.gemm_bf16(a, w, out, ops.EPI_RESID_F32, bias=bias, resid=resid)
    ref = resid + a.float() @ w.float().t() + bias
    assert (out - ref).abs().max().item() < 5e-3
    # in place on the residual stream (how the transformer block uses it)
    x = resid.clone()
    ops.gemm_bf16(a, w, x, ops.EPI_RESID_F32, bias=bias, resid=x)
    assert (x - ref).abs().max().item() < 5e-3


@pytest.mark.parametrize("n,HW,C,dt", [(4, 16384, 128, torch.float32), (3, 4096, 256, torch.bfloat16),
                                       (8, 256, 32, torch.float32), (8, 64, 32, torch.bfloat16), (2, 100, 64, torch.float32)])
def test_groupnorm_stats_and_apply(n, HW, C, dt):
    g = torch.Generator().manual_seed(C + HW)
    x = (torch.randn((n, HW, C), generator=g) * 2 + 0.7).to(DEV).to(dt)
    gamma, beta = torch.randn((C,), generator=g).to(DEV), torch.randn((C,), generator=g).to(DEV)
    sums = torch.empty((n, 32, 3), dtype=torch.float64, device=DEV)
    ops.groupnorm_stats(x, sums, n, HW, C)
    xg = x.double().reshape(n, HW, 32, C // 32)
    acc = fixed_point_sums(sums, n * 32 * 2).reshape(n, 32, 2)
    assert torch.allclose(acc[..., 0], xg.sum((1, 3)), rtol=1e-5, atol=1e-2)
    assert torch.allclose(acc[..., 1], (xg * xg).sum((1, 3)), rtol=1e-5, atol=1e-2)
    x_nchw = x.float().permute(0, 2, 1).reshape(n, C, HW, 1)
    gn = F.group_norm(x_nchw, 32, gamma, beta, eps=1e-6)
    out = torch.empty((n * HW, C), dtype=torch.bfloat16, device=DEV)
    ops.groupnorm_silu_bf16(x, sums, gamma, beta, out, n, HW, C)
    ref = F.silu(gn).reshape(n, C, HW).permute(0, 2, 1).reshape(n * HW, C)
    assert bf16_close(out, ref, 1e-2) and rel_err(out, ref) < 6e-3
    # FiLM: per-image f32 part + per-pixel bf16 part through an image map (-1 = masked)
    mod_img = torch.randn((n, 4 * C + 8), generator=g).to(DEV) * 0.3
    n_src = 2
    mod_pix = (torch.randn((n_src, HW, 2 * C), generator=g) * 0.3).to(DEV).to(torch.bfloat16)
    img_map = torch.tensor([(-1 if i % 3 == 0 else i % n_src) for i in range(n)], dtype=torch.int32, device=DEV)
    sc0, sh0 = 8, 8 + 2 * C
    ops.groupnorm_silu_bf16(x, sums, gamma, beta, out, n, HW, C, mod_img=mod_img, scale_col=sc0, shift_col=sh0,
                            mod_pix=mod_pix, img_map=img_map)
    scale = mod_img[:, None, sc0:sc0 + C].expand(n, HW, C).clone()
    shift = mod_img[:, None, sh0:sh0 + C].expand(n, HW, C).clone()
    for i in range(n):
        if img_map[i] >= 0:
            scale[i] += mod_pix[img_map[i], :, :C].float()
            shift[i] += mod_pix[img_map[i], :, C:].float()
    gn_cl = gn.reshape(n, C, HW).permute(0, 2, 1)
    ref = F.silu(gn_cl * (1 + scale) + shift).reshape(n * HW, C)
    assert bf16_close(out, ref, 1e-2) and rel_err(out, ref) < 6e-3


@pytest.mark.parametrize("M,D,P", [(2048, 576, 256), (1024, 1152, 64), (96, 64, 16), (64, 128, 4)])
def test_rmsnorm_film(M, D, P):
    g = torch.Generator().manual_seed(D)
    x = (torch.randn((M, D), generator=g) * 3).to(DEV)
    w = torch.randn((D,), generator=g).to(DEV)
    n_img = M // P
    mod_img = (torch.randn((n_img, 5 * D), generator=g) * 0.3).to(DEV)
    mod_pix = (torch.randn((2, P, 2 * D), generator=g) * 0.3).to(DEV).to(torch.bfloat16)
    img_map = torch.tensor([(-1 if i % 2 == 0 else (i // 2) % 2) for i in range(n_img)], dtype=torch.int32, device=DEV)
    out = torch.empty((M, D), dtype=torch.bfloat16, device=DEV)
    sc0, sh0 = D, 3 * D
    ops.rmsnorm_film_bf16(x, w, mod_img, sc0, sh0, P, out, mod_pix=mod_pix, img_map=img_map)
    scale = mod_img[:, None, sc0:sc0 + D].expand(n_img, P, D).clone()
    shift = mod_img[:, None, sh0:sh0 + D].expand(n_img, P, D).clone()
    for i in range(n_img):
        if img_map[i] >= 0:
            scale[i] += mod_pix[img_map[i], :, :D].float()
            shift[i] += mod_pix[img_map[i], :, D:].float()
    xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * w
    ref = xn * (1 + scale.reshape(M, D)) + shift.reshape(M, D)
    assert rel_err(out, ref) < 4e-3
    ops.rmsnorm_film_bf16(x, w, mod_img, sc0, sh0, P, out)
    ref = xn * (1 + mod_img[:, sc0:sc0 + D].repeat_interleave(P, 0)) + mod_img[:, sh0:sh0 + D].repeat_interleave(P, 0)
    assert rel_err(out, ref) < 4e-3


@pytest.mark.parametrize("heads,dh,T,gh", [(9, 64, 8, 8), (9, 128, 4, 4), (1, 64, 4, 4), (1, 128, 4, 2), (12, 64, 4, 4),
                                           (20, 64, 2, 2), (5, 128, 3, 2), (16, 128, 2, 2)])
def test_qk_norm_rope(heads, dh, T, gh):
    D, R = heads * dh, 2
    Ntok = T * gh * gh
    M = R * Ntok
    g = torch.Generator().manual_seed(dh + heads)
    buf = torch.randn((M, 3 * D + 4 * D), generator=g).to(DEV).to(torch.bfloat16)    # [q k v | mlp_h], row stride 7D
    qkv = buf[:, : 3 * D]
    orig = qkv.float().clone()
    qw, kw = torch.randn((dh,), generator=g).to(DEV), torch.randn((dh,), generator=g).to(DEV)
    table = rope_cos_sin_table(dh, (T, gh, gh)).to(DEV)
    scale = 1.4426950408889634 / math.sqrt(dh)
    ops.qk_norm_rope(qkv, qw, kw, table, Ntok, heads, dh, scale)
    q, k, v = orig.reshape(M, 3, heads, dh).unbind(1)
    rms = lambda t, w: t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6) * w

    def rope(t):
        cs = table[torch.arange(M, device=DEV) % Ntok][:, None]          # [M, 1, dh/2, 2]
        x0, x1 = t[..., 0::2], t[..., 1::2]
        return torch.stack([x0 * cs[..., 0] - x1 * cs[..., 1], x1 * cs[..., 0] + x0 * cs[..., 1]], -1).flatten(-2)
    ref = torch.stack([rope(rms(q, qw)) * scale, rope(rms(k, kw)), v], 1).reshape(M, 3 * D)
    assert bf16_close(qkv, ref, 1e-2) and rel_err(qkv, ref) < 5e-3
    assert torch.equal(buf[:, 3 * D:].float(), buf[:, 3 * D:].float())


@pytest.mark.parametrize("heads,dh,T,gh,R", [(9, 64, 8, 8, 2), (9, 128, 4, 4, 2), (1, 64, 4, 4, 3), (2, 128, 4, 2, 1),
                                             (3, 64, 5, 3, 1)])
def test_gemm_qknorm_rope_epilogue(heads, dh, T, gh, R):
    """QKV GEMM with fused q/k RMSNorm(head_dim) + RoPE-3D + q scale (incl. a token count that is not a multiple of 32)."""
    D = heads * dh
    Ntok = T * gh * gh
    M = R * Ntok
    g = torch.Generator().manual_seed(dh + heads + T)
    a = torch.randn((M, D), generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn((3 * D, D), generator=g) / math.sqrt(D)).to(DEV).to(torch.bfloat16)
    bias = torch.randn((3 * D,), generator=g).to(DEV)
    qw, kw = torch.randn((dh,), generator=g).to(DEV), torch.randn((dh,), generator=g).to(DEV)
    table = rope_cos_sin_table(dh, (T, gh, gh)).to(DEV)
    scale = 1.4426950408889634 / math.sqrt(dh)
    buf = torch.full((M, 3 * D), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm_bf16(a, w, buf, ops.EPI_QKNORM_ROPE_BF16, bias=bias, rope_cs=table, tokens_per_sample=Ntok, model_dim=D,
                  head_dim=dh, q_scale=scale, qn_w=qw, kn_w=kw)
    acc = a.float() @ w.float().t() + bias
    q, k, v = acc.reshape(M, 3, heads, dh).unbind(1)
    rms = lambda t, wn: t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6) * wn

    def rope(t):
        cs = table[torch.arange(M, device=DEV) % Ntok][:, None]
        x0, x1 = t[..., 0::2], t[..., 1::2]
        return torch.stack([x0 * cs[..., 0] - x1 * cs[..., 1], x1 * cs[..., 0] + x0 * cs[..., 1]], -1).flatten(-2)
    ref = torch.stack([rope(rms(q, qw)) * scale, rope(rms(k, kw)), v], 1).reshape(M, 3 * D)
    assert torch.isfinite(buf.float()).all()
    assert bf16_close(buf, ref, 1e-2) and rel_err(buf, ref) < 5e-3


def test_pool_upsample_sub():
    g = torch.Generator().manual_seed(3)
    n, H, W, C = 3, 16, 32, 64
    x = torch.randn((n, H, W, C), generator=g).to(DEV)
    ref = F.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    o16 = torch.empty((n, H // 2, W // 2, C), dtype=torch.bfloat16, device=DEV)
    ops.avgpool2x2(x, o16, n, H, W, C)
    assert (o16.float() - ref).abs().max().item() < 2e-2
    o32 = torch.empty((n, H // 2, W // 2, C), device=DEV)
    ops.avgpool2x2(x, o32, n, H, W, C)
    assert (o32 - ref).abs().max().item() < 1e-6
    xb = x.to(torch.bfloat16)
    ops.avgpool2x2(xb, o16, n, H, W, C)
    assert (o16.float() - F.avg_pool2d(xb.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)).abs().max().item() < 2e-2
    skip = torch.randn((n, H, W, C), generator=g).to(DEV)
    low = torch.randn((n, H // 2, W // 2, C), generator=g).to(DEV)
    out = torch.empty_like(skip)
    ops.upsample2x_add(low, skip, out, n, H, W, C)
    ref = F.interpolate(low.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1) + skip
    assert torch.equal(out, ref)
    d16 = torch.empty((n, H, W, C), dtype=torch.bfloat16, device=DEV)
    ops.sub_bf16(x, skip, d16)
    assert torch.equal(d16, (x - skip).to(torch.bfloat16))


@pytest.mark.parametrize("res,p,B,T", [(32, 2, 2, 4), (64, 2, 1, 3)])
def test_pose_ray_patches_vs_oracle(res, p, B, T):
    from oracle.cases import synthetic_poses
    from oracle.pose import ray_encoding
    from dfot_b200.algorithms.dfot.dfot_video_pose import camera_table, ray_freq_scale
    poses = synthetic_poses(B, T)
    enc = ray_encoding(poses, res, "first", None, "ray_encoding")           # (B, T, 180, res, res) fp32, CPU oracle
    cams = camera_table(poses, res, "first", None).to(DEV)
    g = res // p
    out = torch.empty((B * T * g * g, p * p * 180), dtype=torch.bfloat16, device=DEV)
    ops.pose_ray_patches(cams.reshape(B * T, 16).contiguous(), ray_freq_scale().to(DEV), out, B * T, res, p)
    ref = enc.reshape(B * T, 180, g, p, g, p).permute(0, 2, 4, 3, 5, 1).reshape(B * T * g * g, p * p * 180)
    err = (out.float().cpu() - ref).abs()
    # the top octaves (2^13, 2^14 * pi) amplify 1-ulp differences of the ray direction by ~5e4: compare them loosely
    ch = torch.arange(180) % 15
    lo = (ch < 10).repeat(p * p)
    assert err[:, lo].max().item() < 1e-2, err[:, lo].max()
    assert err[:, ~lo].max().item() < 0.25, err[:, ~lo].max()
    # rows with padding (ld > p*p*180) take the element-wise write-out: same values, padding untouched
    wide = torch.full((B * T * g * g, p * p * 180 + 8), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.pose_ray_patches(cams.reshape(B * T, 16).contiguous(), ray_freq_scale().to(DEV), wide[:, : p * p * 180], B * T, res, p)
    assert torch.equal(wide[:, : p * p * 180], out) and bool((wide[:, p * p * 180:] == 7.0).all())
