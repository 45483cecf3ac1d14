"""Test infrastructure: CPU (torch fp32) restatements of the *contracts* of the C-ABI kernels in include/dfot_b200.h,
used to drive the product's host-side orchestration (weight packing, buffer flow, column offsets, layouts) on a machine
without a GPU.  Never imported by the product; the `-m gpu` tests check the real kernels against the same contracts."""
import math

import torch
import torch.nn.functional as F

from dfot_b200 import ops

BF = torch.bfloat16


def _col_offset(t):
    """Element offset of a view's first element inside its row (0 for whole tensors): the kernels want 16-byte aligned
    pointers, and the base allocations are, so the offset decides."""
    return t.storage_offset() % t.stride(0) if t.dim() == 2 and t.stride(0) > 0 else 0


def cast_bf16(src, out=None):
    if out is None:
        return src.to(BF)
    out.copy_(src.to(BF))
    return out


def patchify_bf16(x, out, frames, C, H, W, p):
    x = x.float().reshape(frames, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(-1, C * p * p)
    out[:, : C * p * p] = x.to(BF)


def unpatchify(tok, x, frames, C, H, W, p):
    t = tok[:, : p * p * C].reshape(frames, H // p, W // p, p, p, C).permute(0, 5, 1, 3, 2, 4).reshape(frames, C, H, W)
    x.copy_(t.reshape(x.shape).to(x.dtype))


def _epilogue(acc, out, epilogue, bias, resid):
    if bias is not None:
        acc = acc + bias
    if epilogue == ops.EPI_SILU_BF16:
        acc = F.silu(acc)
    elif epilogue == ops.EPI_GELU_BF16:
        acc = F.gelu(acc, approximate="tanh")
    elif epilogue == ops.EPI_RESID_F32:
        acc = acc + resid.reshape(acc.shape)
    elif epilogue not in (ops.EPI_F32, ops.EPI_BF16):
        raise NotImplementedError(epilogue)
    out.copy_(acc.reshape(out.shape).to(out.dtype))


def gemm_bf16(a, w, out, epilogue, bias=None, resid=None, M=None, **kw):
    assert a.dtype == BF and w.dtype == BF
    # the kernel's contract (dfot_gemm_bf16: 16-byte TMA strides, 16-byte aligned operands and epilogue vectors)
    assert w.shape[1] % 8 == 0 and a.stride(0) % 8 == 0 and w.stride(0) % 8 == 0, "gemm: K, lda, ldw must be multiples of 8"
    assert w.shape[0] % 8 == 0 and out.stride(0) % 8 == 0, "gemm: N and ldc must be multiples of 8"
    assert _col_offset(a) % 8 == 0 and _col_offset(w) % 8 == 0, "gemm: A, W must be 16-byte aligned"
    assert _col_offset(out) % (4 if out.dtype == torch.float32 else 8) == 0, "gemm: C must be 16-byte aligned"
    if kw.get("gate") is not None:
        assert _col_offset(kw["gate"]) % 4 == 0 and kw.get("ld_gate", 0) % 4 == 0, "gemm: gate must be 16-byte aligned"
    if epilogue in (ops.EPI_QKV_ROPE_BF16, ops.EPI_QKNORM_ROPE_BF16):
        assert kw["head_dim"] % 2 == 0 and kw["model_dim"] % kw["head_dim"] == 0 and w.shape[0] == 3 * kw["model_dim"]
    if epilogue == ops.EPI_QKNORM_ROPE_BF16:
        assert kw["head_dim"] in (64, 128) and w.shape[0] > 128
    M = a.shape[0] if M is None else M
    acc = a[:M].float() @ w.float().t()
    if epilogue == ops.EPI_QKNORM_ROPE_BF16:      # q/k RMSNorm(head_dim) * weight, RoPE-3D, q * q_scale; v untouched
        acc = acc + bias
        D, dh, tps = kw["model_dim"], kw["head_dim"], kw["tokens_per_sample"]
        heads = D // dh
        q, k, v = (acc[:, i * D:(i + 1) * D].reshape(M, heads, dh) for i in range(3))
        cs = kw["rope_cs"][torch.arange(M) % tps][:, None]

        def f(t, wn, mul):
            t = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + kw.get("qk_eps", 1e-6)) * wn
            x0, x1 = t[..., 0::2], t[..., 1::2]
            return torch.stack([x0 * cs[..., 0] - x1 * cs[..., 1], x1 * cs[..., 0] + x0 * cs[..., 1]], -1).flatten(-2) * mul
        res = torch.cat([f(q, kw["qn_w"], kw["q_scale"]).reshape(M, D), f(k, kw["kn_w"], 1.0).reshape(M, D),
                         v.reshape(M, D)], 1)
        out[:M].copy_(res.to(out.dtype))
        return
    if epilogue == ops.EPI_QKV_ROPE_BF16:         # RoPE-3D on the q and k columns (pairs (2i, 2i+1) per head), q * q_scale
        acc = acc + bias
        D, dh, tps = kw["model_dim"], kw["head_dim"], kw["tokens_per_sample"]
        heads = D // dh
        q, k, v = (acc[:, i * D:(i + 1) * D].reshape(M, heads, dh) for i in range(3))
        cs = kw["rope_cs"][torch.arange(M) % tps][:, None]

        def rot(t, mul):
            x0, x1 = t[..., 0::2], t[..., 1::2]
            return torch.stack([x0 * cs[..., 0] - x1 * cs[..., 1], x1 * cs[..., 0] + x0 * cs[..., 1]], -1).flatten(-2) * mul
        out[:M].copy_(torch.cat([rot(q, kw["q_scale"]).reshape(M, D), rot(k, 1.0).reshape(M, D), v.reshape(M, D)], 1).to(out.dtype))
        return
    if epilogue == ops.EPI_GATE_RESID_F32:        # out = resid + gate[frame(m), n] * (acc + bias)
        frame = torch.arange(M) // kw["tokens_per_frame"]
        out[:M].copy_(resid[:M] + kw["gate"][frame][:, : acc.shape[1]] * (acc + bias))
        return
    if epilogue == ops.EPI_GATE_LNRESID_F32:      # the residual base rebuilt from x = resid and the row statistics
        N = acc.shape[1]
        frame = torch.arange(M) // kw["tokens_per_frame"]
        st = kw["ln_stats"].reshape(-1, 2)[:M]
        base = ((resid[:M] - st[:, :1]) * st[:, 1:]) * (1 + kw["ln_scale"][frame][:, :N]) + kw["ln_shift"][frame][:, :N]
        out[:M].copy_(base + kw["gate"][frame][:, :N] * (acc + bias))
        return
    _epilogue(acc, out[:M], epilogue, bias, None if resid is None else resid[:M])


def adaln_layernorm(x, mod, shift_col, scale_col, tokens_per_frame, y_f32=None, y_bf16=None, eps=1e-6, stats=None):
    M, D = x.shape
    assert D % 4 == 0 and mod.shape[-1] % 4 == 0 and shift_col % 4 == 0 and scale_col % 4 == 0 and D <= 4096, "adaln contract"
    frame = torch.arange(M) // tokens_per_frame
    mean = x.float().mean(-1, keepdim=True)
    rstd = torch.rsqrt(((x.float() - mean) ** 2).mean(-1, keepdim=True) + eps)
    y = ((x.float() - mean) * rstd) * (1 + mod[frame, scale_col:scale_col + D]) + mod[frame, shift_col:shift_col + D]
    if stats is not None:
        stats.reshape(-1, 2).copy_(torch.cat([mean, rstd], 1))
    if y_f32 is not None:
        y_f32.copy_(y)
    if y_bf16 is not None:
        y_bf16.copy_(y.to(BF))


def patch_mix_bf16(y, u, out, R, L, P, Mc):
    D = y.shape[-1]
    assert D % 4 == 0, "patch_mix: D % 4"
    s = torch.einsum("nc,rlnd->rcld", u.reshape(P, Mc).float(), y.reshape(R, L, P, D).float())
    out.copy_(s.reshape(out.shape).to(BF))


def patch_expand_gate_resid(x, y, z, pu, pb, gate, ld_gate, R, L, P, Mc):
    D = x.shape[-1]
    assert D % 4 == 0 and (gate is None or (ld_gate % 4 == 0 and ld_gate >= D and _col_offset(gate) % 4 == 0)), "patch_expand contract"
    assert y is None or x.data_ptr() != y.data_ptr(), "patch_expand: x must not alias y"
    s = torch.einsum("cn,rcld->rlnd", pu.reshape(Mc, P).float(), z.reshape(R, Mc, L, D).float())
    if pb is not None:
        s = s + pb.reshape(1, 1, P, D)
    if gate is not None:
        s = gate[:, :D].reshape(R, L, 1, D) * s
    if y is not None:
        s = y.reshape(R, L, P, D) + s
    x.copy_(s.reshape(x.shape))


def gemm_bf16_splitk(a, w, parts, splits, M=None):
    M = a.shape[0] if M is None else M
    N, K = w.shape
    kb = -(-K // 64)
    out = parts.reshape(-1)[: M * splits * N].view(M, splits, N)
    for s in range(splits):
        k0, k1 = 64 * (s * kb // splits), min(K, 64 * ((s + 1) * kb // splits))
        out[:, s] = a[:M, k0:k1].float() @ w[:, k0:k1].float().t()


def splitk_gate_resid_adaln(parts, splits, bias, resid, mod, gate_col, shift_col, scale_col, tokens_per_frame, x_out=None,
                            y_f32=None, y_bf16=None, eps=1e-6):
    M, D = resid.shape
    acc = parts.reshape(-1)[: M * splits * D].view(M, splits, D).sum(1)
    if bias is not None:
        acc = acc + bias
    frame = torch.arange(M) // tokens_per_frame
    x = resid + mod[frame, gate_col:gate_col + D] * acc
    if x_out is not None:
        x_out.copy_(x)
    if shift_col >= 0:
        mean = x.mean(-1, keepdim=True)                       # (the arithmetic of adaln_layernorm above)
        rstd = torch.rsqrt(((x - mean) ** 2).mean(-1, keepdim=True) + eps)
        y = ((x - mean) * rstd) * (1 + mod[frame, scale_col:scale_col + D]) + mod[frame, shift_col:shift_col + D]
        if y_f32 is not None:
            y_f32.copy_(y)
        if y_bf16 is not None:
            y_bf16.copy_(y.to(BF))


def silu_sum_bf16(a, b, row_mask, rows_per_mask, out):
    s = a.float()
    if b is not None:
        keep = 1.0
        if row_mask is not None:
            keep = (row_mask.reshape(-1) == 0).float().repeat_interleave(rows_per_mask)[:, None]
        s = s + b.float() * keep
    out.copy_(F.silu(s).to(BF))


def conv3x3_bf16(x, w, out, epilogue, bias=None, resid=None, gn_sums=None, gn_groups=32, gn_eps=1e-6):
    assert x.dtype == BF and w.dtype == BF
    _conv_contract(x.shape[1], x.shape[2], x.shape[3], w.shape[0], out.stride(0) if out.dim() == 2 else w.shape[0])
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), None, padding=1).permute(0, 2, 3, 1)
    _epilogue(y.reshape(-1, w.shape[0]), out, epilogue, bias, resid)
    if gn_sums is not None:      # side output: statistics of the stored output
        n, H, W_, _ = x.shape
        groupnorm_stats(out, gn_sums, n, H * W_, w.shape[0], gn_groups, gn_eps)


def groupnorm_stats(x, sums, n_img, HW, C, groups=32, eps=1e-6):
    xg = x.double().reshape(n_img, HW, groups, C // groups)
    cnt = HW * (C // groups)
    mean = xg.sum((1, 3)) / cnt
    var = ((xg * xg).sum((1, 3)) / cnt - mean * mean).clamp(min=0)
    sums[..., 0] = mean                      # emulation keeps (mean, rstd) in the first two slots of the workspace
    sums[..., 1] = torch.rsqrt(var.float() + eps).double()


def _film(n_img, HW, C, mod_img, scale_col, shift_col, mod_pix, img_map):
    scale = mod_img[:, None, scale_col:scale_col + C].expand(n_img, HW, C).clone()
    shift = mod_img[:, None, shift_col:shift_col + C].expand(n_img, HW, C).clone()
    if mod_pix is not None:
        pix = mod_pix.float().reshape(-1, HW, 2 * C)
        for i, src in enumerate(img_map.tolist()):
            if src >= 0:
                scale[i] += pix[src, :, :C]
                shift[i] += pix[src, :, C:]
    return scale, shift


def groupnorm_silu_bf16(x, sums, gamma, beta, out, n_img, HW, C, groups=32, mod_img=None, scale_col=0,
                        shift_col=0, mod_pix=None, img_map=None):
    xg = x.float().reshape(n_img, HW, groups, C // groups)
    y = (xg - sums[..., 0][:, None, :, None].float()) * sums[..., 1][:, None, :, None].float()
    y = y.reshape(n_img, HW, C) * gamma + beta
    if mod_img is not None:
        scale, shift = _film(n_img, HW, C, mod_img, scale_col, shift_col, mod_pix, img_map)
        y = y * (1 + scale) + shift
    out.copy_(F.silu(y).reshape(out.shape).to(BF))


def rmsnorm_film_bf16(x, weight, mod_img, scale_col, shift_col, tokens_per_img, out, mod_pix=None, img_map=None,
                      eps=1e-6):
    M, D = x.shape
    n_img = M // tokens_per_img
    scale, shift = _film(n_img, tokens_per_img, D, mod_img, scale_col, shift_col, mod_pix, img_map)
    xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight
    out.copy_((xn * (1 + scale.reshape(M, D)) + shift.reshape(M, D)).to(BF))


def qk_norm_rope(qkv, q_weight, k_weight, rope_cs, tokens_per_sample, heads, head_dim, q_scale, eps=1e-6):
    assert head_dim in (64, 128) and 2 * heads * head_dim <= 4096 and qkv.stride(0) % 8 == 0, "qk_norm_rope contract"
    M = qkv.shape[0]
    D = heads * head_dim
    q, k = (qkv[:, i * D:(i + 1) * D].float().reshape(M, heads, head_dim) for i in range(2))
    cs = rope_cs[torch.arange(M) % tokens_per_sample][:, None]

    def f(t, w, mul):
        t = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + eps) * w
        x0, x1 = t[..., 0::2], t[..., 1::2]
        return (torch.stack([x0 * cs[..., 0] - x1 * cs[..., 1], x1 * cs[..., 0] + x0 * cs[..., 1]], -1).flatten(-2)
                * mul).reshape(M, D)
    qkv[:, :D] = f(q, q_weight, q_scale).to(BF)
    qkv[:, D:2 * D] = f(k, k_weight, 1.0).to(BF)


def attention(qkv, out, R, Ntok, heads, head_dim, score_bound=0.0):
    assert head_dim in (64, 72, 128), f"attention: head_dim {head_dim} unsupported (64, 72, 128)"      # the kernel's contract
    assert out.stride(0) % 8 == 0 and out.stride(0) >= heads * head_dim and _col_offset(out) % 8 == 0, "attention: ld_out / alignment"
    D = heads * head_dim
    q, k, v = qkv.float().reshape(R, Ntok, 3, heads, head_dim).permute(2, 0, 3, 1, 4).unbind(0)
    w = torch.softmax(q @ k.transpose(-1, -2) * math.log(2.0), dim=-1)       # q carries scale * log2(e)
    out.copy_((w @ v).transpose(1, 2).reshape(R * Ntok, D).to(BF))


def avgpool2x2(x, out, n_img, H, W, C):
    y = F.avg_pool2d(x.float().reshape(n_img, H, W, C).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    out.copy_(y.reshape(out.shape).to(out.dtype))


def sub_bf16(a, b, out):
    out.copy_((a - b).to(BF))


def upsample2x_add(low, skip, out, n_img, H, W, C):
    up = F.interpolate(low.reshape(n_img, H // 2, W // 2, C).permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    out.copy_((up.permute(0, 2, 3, 1).reshape(skip.shape) + skip))


def pose_ray_patches(cams, freq_scale, out, frames, res, p):
    lin = torch.arange(res, dtype=torch.float32) + 0.5
    fx, fy, px, py = (cams[:, i][:, None, None] for i in range(4))
    dx = ((lin[None, None, :] - px) / fx).expand(frames, res, res)
    dy = ((lin[None, :, None] - py) / fy).expand(frames, res, res)
    Rinv = cams[:, 4:13].reshape(frames, 3, 3)
    d = torch.stack([Rinv[:, i, 0][:, None, None] * dx + Rinv[:, i, 1][:, None, None] * dy + Rinv[:, i, 2][:, None, None]
                     for i in range(3)], -1)
    o = cams[:, None, None, 13:16].expand(frames, res, res, 3)

    def enc(v):
        e = (v[..., None] * freq_scale).flatten(-2)
        return torch.sin(torch.cat([e, e + 0.5 * math.pi], -1))
    full = torch.cat([enc(o), enc(d)], -1)                                      # frames, res, res, 12*n_freq
    g, ch = res // p, full.shape[-1]
    rows = full.reshape(frames, g, p, g, p, ch).permute(0, 1, 3, 2, 4, 5).reshape(frames * g * g, p * p * ch)
    out[:, : p * p * ch] = rows.to(BF)


def noise_features(levels, out, fourier_freqs=None, fourier_phases=None):
    k = levels.float().reshape(-1, 1)
    dim = out.shape[-1]
    if fourier_freqs is not None:
        v = torch.cos(k * fourier_freqs + fourier_phases) * math.sqrt(2.0)
    else:
        half = dim // 2
        f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
        v = torch.cat([torch.cos(k * f), torch.sin(k * f)], -1)
    out.copy_(v.reshape(out.shape).to(BF))


# ------------------------------------------------------------------ VAE-decode row (clips [B, 2 + T, H, W, C])
def _conv_contract(H, W, Cin, Cout, ldc):
    """dfot_conv3x3_bf16 / dfot_conv3d_causal_bf16: channel counts and the image-tile rules of the 4-D TMA box."""
    assert Cin % 8 == 0 and Cout % 8 == 0 and ldc % 8 == 0, "conv3x3: Cin, Cout, ldc must be multiples of 8"
    pow2 = lambda n: n > 0 and n & (n - 1) == 0
    if W >= 128:
        assert W % 128 == 0, "conv3x3: W must be a multiple of 128 or a power of two"
    else:
        assert pow2(W), "conv3x3: W must be a power of two below 128"
        assert H % (128 // W) == 0 if H * W >= 128 else pow2(H), "conv3x3: H vs the 128-pixel tile"


def conv3d_causal_bf16(x, w, out, epilogue, bias=None, resid=None):
    assert x.dtype == BF and w.dtype == BF
    _conv_contract(x.shape[1], x.shape[2], x.shape[3], w.shape[0], out.stride(0) if out.dim() == 2 else w.shape[0])
    kt = w.shape[1]
    xs = x.float().permute(3, 0, 1, 2)[None]                                    # [1, Cin, n_in, H, W]
    y = F.conv3d(xs, w.float().permute(0, 4, 1, 2, 3), None, padding=(0, 1, 1))[0].permute(1, 2, 3, 0)
    assert y.shape[0] == x.shape[0] - kt + 1
    _epilogue(y.reshape(-1, w.shape[0]), out, epilogue, bias, resid)


def _strided_images(x, n_img, HW, img_stride, C):
    return torch.as_strided(x, (n_img, HW, C), (img_stride, C, 1))


def groupnorm_stats_strided(x, sums, n_img, HW, img_stride, C, groups=32, eps=1e-6):
    groupnorm_stats(_strided_images(x, n_img, HW, img_stride, C), sums, n_img, HW, C, groups, eps)


def groupnorm_apply_bf16(x, sums, gamma, beta, out, n_img, HW, img_stride, C, groups=32, silu=True):
    xg = _strided_images(x, n_img, HW, img_stride, C).float().reshape(n_img, HW, groups, C // groups)
    y = ((xg - sums[..., 0].float()[:, None, :, None]) * sums[..., 1].float()[:, None, :, None]).reshape(n_img, HW, C)
    y = y * gamma + beta
    if silu:
        y = F.silu(y)
    _strided_images(out, n_img, HW, img_stride, C).copy_(y.to(BF))


def vae_fill_pad_frames(x, B, T, frame_elems):
    v = x.view(B, 2 + T, frame_elems)
    v[:, :2] = v[:, 2:3]


def vae_upsample2x_bf16(x, out, B, T_in, H, W, C, temporal):
    v = x.view(B, 2 + T_in, H, W, C)[:, 2:].float().permute(0, 4, 1, 2, 3)         # b c t h w
    if not temporal:
        y = F.interpolate(v.reshape(B, C * T_in, H, W), scale_factor=(2, 2), mode="nearest").reshape(B, C, T_in, 2 * H, 2 * W)
    else:
        y = F.interpolate(v[:, :, :1], scale_factor=(1, 2, 2), mode="trilinear")
        if T_in > 1:
            y = torch.cat([y, F.interpolate(v[:, :, 1:], scale_factor=(2, 2, 2), mode="trilinear")], 2)
    o = out.view(B, 2 + y.shape[2], 2 * H, 2 * W, C)
    o[:, 2:] = y.permute(0, 2, 3, 4, 1).to(BF)
    o[:, :2] = o[:, 2:3]


def upsample2x_nearest_bf16(x, out, n_img, H, W, C):
    y = F.interpolate(x.view(n_img, H, W, C).float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    out.view(n_img, 2 * H, 2 * W, C).copy_(y.permute(0, 2, 3, 1).to(BF))


def softmax_rows_bf16(s, p, scale=1.0):
    p.copy_(torch.softmax(s.float() * scale, -1).to(BF))


# ---------------------------------------------------------------- DC-AE decoder glue (csrc/dcae.cu)
def relu_bf16(x):
    x.copy_(F.relu(x.float()).to(BF))


def pixel_shuffle2x(conv, C, n_img, H, W, shortcut=None, repeats=1, out_f32=None, out_bf16=None):
    y = conv[:, : 4 * C].reshape(n_img, H, W, 4 * C).float()
    if shortcut is not None:
        y = y + shortcut.reshape(n_img, H, W, -1).repeat_interleave(repeats, dim=-1)
    y = F.pixel_shuffle(y.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)          # [n, 2H, 2W, C]
    if out_f32 is not None:
        out_f32.copy_(y.reshape(out_f32.shape))
    if out_bf16 is not None:
        out_bf16.copy_(y.reshape(out_bf16.shape).to(BF))


def linear_attention_relu(qkv, out, n_img, HW, heads, head_dim, eps=1e-15):
    d = head_dim
    t = qkv[:, : 3 * heads * d].float().reshape(n_img, HW, heads, 3 * d)
    q, k, v = F.relu(t[..., :d]), F.relu(t[..., d:2 * d]), t[..., 2 * d:]
    v1 = torch.cat([v, torch.ones_like(v[..., :1])], -1)
    kv = torch.einsum("ntha,nthb->nhab", k, v1)
    o = torch.einsum("ntha,nhab->nthb", q, kv)
    out[:, : heads * d] = (o[..., :d] / (o[..., d:] + eps)).reshape(n_img * HW, heads * d)


def dwconv3x3_glu_bf16(x, w, b, out, n_img, H, W, Ch):
    xi = x.float().reshape(n_img, H, W, 2 * Ch).permute(0, 3, 1, 2)
    y = F.conv2d(xi, w.reshape(2 * Ch, 1, 3, 3), b, padding=1, groups=2 * Ch)
    h, gate = torch.chunk(y, 2, dim=1)
    out.copy_((h * F.silu(gate)).permute(0, 2, 3, 1).reshape(out.shape).to(BF))


def rmsnorm_affine(x, w, b, eps, resid=None, relu=False, out_f32=None, out_bf16=None):
    y = x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + eps) * w + b
    if resid is not None:
        y = y + resid
    if relu:
        y = F.relu(y)
    if out_f32 is not None:
        out_f32.copy_(y)
    if out_bf16 is not None:
        out_bf16.copy_(y.to(BF))


ALL = ["relu_bf16", "pixel_shuffle2x", "linear_attention_relu", "dwconv3x3_glu_bf16", "rmsnorm_affine", "cast_bf16", "patchify_bf16", "unpatchify", "gemm_bf16", "conv3x3_bf16", "groupnorm_stats", "groupnorm_silu_bf16",
       "rmsnorm_film_bf16", "qk_norm_rope", "attention", "avgpool2x2", "sub_bf16", "upsample2x_add", "pose_ray_patches",
       "noise_features", "conv3d_causal_bf16", "groupnorm_stats_strided", "groupnorm_apply_bf16",
       "vae_upsample2x_bf16", "upsample2x_nearest_bf16", "vae_fill_pad_frames", "softmax_rows_bf16", "adaln_layernorm",
       "silu_sum_bf16", "patch_mix_bf16", "patch_expand_gate_resid", "gemm_bf16_splitk",
       "splitk_gate_resid_adaln"]


def install(monkeypatch):
    """Route the product's op calls to the CPU contract emulations and lift its CUDA-only guards (tests only)."""
    g = globals()
    for name in ALL:
        monkeypatch.setattr(ops, name, g[name])
    monkeypatch.setattr(ops, "require_cuda", lambda *a, **k: None)


def install_raw():
    """Same as install() without a pytest fixture (sub-processes); returns a function that undoes it."""
    g = globals()
    saved = {name: getattr(ops, name) for name in ALL + ["require_cuda"]}
    for name in ALL:
        setattr(ops, name, g[name])
    ops.require_cuda = lambda *a, **k: None

    def restore():
        for name, fn in saved.items():
            setattr(ops, name, fn)
    return restore
