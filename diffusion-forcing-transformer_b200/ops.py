"""Tensor-level wrappers over the C ABI (torch is used for device memory and streams only).

Every wrapper checks device / dtype / contiguity, launches on torch's current CUDA stream and
raises RuntimeError with the library's message on failure.  No op has a CPU implementation.
"""
import ctypes
from typing import Optional

import torch

from . import _abi
from ._abi import (BF16, EPI_BF16, EPI_F32, EPI_GATE_RESID_F32, EPI_GELU_BF16, EPI_QKV_ROPE_BF16,  # noqa: F401
                   EPI_QKNORM_ROPE_BF16, EPI_RESID_F32, EPI_SILU_BF16, EPI_GATE_LNRESID_F32, F32, I64)

_DTYPE_TAG = {torch.float32: F32, torch.bfloat16: BF16, torch.int64: I64}


_replayed_launches = 0


def count_replayed_launches(n: int) -> None:
    """Kernels of this library that ran through a CUDA-graph replay (they bypass the C-ABI launch counter)."""
    global _replayed_launches
    _replayed_launches += n


def total_launches() -> int:
    """Kernels of this library launched so far: direct C-ABI launches + kernels inside replayed graphs
    (the capture itself is counted once by the C-ABI counter and executes nothing)."""
    return _abi.launch_count() + _replayed_launches


def require_cuda(device, what: str) -> None:
    """There is no CPU implementation of anything in this package: callers guard their entry points with this."""
    if torch.device(device).type != "cuda":
        raise RuntimeError(f"dfot_b200: {what} runs on CUDA only (no CPU fallback); move it to a B200")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"dfot_b200: `{name}` must be a CUDA tensor (no CPU fallback exists)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"dfot_b200: `{name}` must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"dfot_b200: `{name}` must be contiguous")


def sampler_step_hg(x, model_out, model_in_next, upd, prep, noise_ddim, noise_hist, noise_excl, B, nfe, T,
                    max_noise_row=None):
    """K4. x [B,T,...] f32 in place; model_out / model_in_next [B*nfe,T,...] f32|bf16 or None;
    upd / prep: uint8 CUDA tensors holding packed dfot_frame_update / dfot_frame_prepare records; `max_noise_row`: the
    largest `noise_row` of the prepare table (known to the host planner), checked against the rows of `noise_hist`."""
    _need(x, torch.float32, "x")
    F = x[0, 0].numel()
    for t, n in ((model_out, "model_out"), (model_in_next, "model_in_next")):
        if t is not None:
            _need(t, None, n)
            if t.dtype not in (torch.float32, torch.bfloat16) or t.numel() != B * nfe * T * F:
                raise RuntimeError(f"dfot_b200: `{n}` has wrong dtype/size")
    if x.dim() < 3 or tuple(x.shape[:2]) != (B, T):
        raise RuntimeError(f"dfot_b200: `x` must be [B={B}, T={T}, ...], got {tuple(x.shape)}")
    for t, n in ((noise_ddim, "noise_ddim"), (noise_excl, "noise_excl")):
        if t is not None:
            _need(t, torch.float32, n)
            if t.numel() != B * nfe * T * F:
                raise RuntimeError(f"dfot_b200: `{n}` has {t.numel()} elements, expected B*nfe*T*F = {B * nfe * T * F}")
    if noise_hist is not None:
        _need(noise_hist, torch.float32, "noise_hist")
        if noise_hist.numel() % (T * F):
            raise RuntimeError(f"dfot_b200: `noise_hist` must hold whole [T, F] rows, got {noise_hist.numel()} elements")
        if prep is not None and max_noise_row is not None and noise_hist.numel() < (max_noise_row + 1) * T * F:
            raise RuntimeError(f"dfot_b200: `noise_hist` holds {noise_hist.numel() // (T * F)} rows, the prepare table "
                               f"addresses row {max_noise_row}")
    for t, n, sz in ((upd, "upd", 24), (prep, "prep", 16)):
        if t is not None:
            _need(t, torch.uint8, n)
            if t.numel() != B * nfe * T * sz:
                raise RuntimeError(f"dfot_b200: `{n}` table has {t.numel()} bytes, expected {B * nfe * T * sz}")
    rc = _abi.lib().dfot_sampler_step_hg(
        x.data_ptr(), _ptr(model_out), _DTYPE_TAG[model_out.dtype] if model_out is not None else F32,
        _ptr(model_in_next), _DTYPE_TAG[model_in_next.dtype] if model_in_next is not None else F32,
        _ptr(upd), _ptr(prep), _ptr(noise_ddim), _ptr(noise_hist), _ptr(noise_excl), B, nfe, T, F, _stream())
    _abi.check(rc, "sampler_step_hg")


def adaln_layernorm(x, mod, shift_col, scale_col, tokens_per_frame, y_f32=None, y_bf16=None, eps=1e-6, stats=None):
    """K1. x [M,D] f32, mod [frames, mod_ld] f32; stats [M,2] f32 (optional side output: mean, rstd per row)."""
    _need(x, torch.float32, "x")
    _need(mod, torch.float32, "mod")
    M, D = x.shape
    if y_f32 is not None:
        _need(y_f32, torch.float32, "y_f32")
    if y_bf16 is not None:
        _need(y_bf16, torch.bfloat16, "y_bf16")
    if stats is not None:
        _need(stats, torch.float32, "stats")
        if stats.numel() != 2 * M:
            raise RuntimeError("dfot_b200: `stats` must hold (mean, rstd) of every row: [M, 2] f32")
    rc = _abi.lib().dfot_adaln_layernorm_stats(x.data_ptr(), mod.data_ptr(), mod.shape[-1], shift_col, scale_col,
                                               _ptr(y_f32), _ptr(y_bf16), _ptr(stats), M, D, tokens_per_frame, eps, _stream())
    _abi.check(rc, "adaln_layernorm")


def _gn_side_output(e, gn_sums, gn_rows_per_img, gn_groups, gn_eps, M, N):
    """GroupNorm statistics of the output as a side product of the epilogue (see include/dfot_b200.h)."""
    if gn_sums is None:
        return
    _need(gn_sums, torch.float64, "gn_sums")
    if gn_rows_per_img <= 0 or M % gn_rows_per_img or gn_sums.numel() != 3 * (M // gn_rows_per_img) * gn_groups:
        raise RuntimeError("dfot_b200: gn_sums must hold 3 * n_img * groups doubles (n_img = M / gn_rows_per_img)")
    e.gn_sums, e.gn_rows_per_img, e.gn_groups, e.gn_eps = gn_sums.data_ptr(), gn_rows_per_img, gn_groups, gn_eps


def gemm_bf16(a, w, out, epilogue, bias=None, resid=None, gate=None, ld_gate=0, tokens_per_frame=1, rope_cs=None,
              tokens_per_sample=1, model_dim=0, head_dim=0, q_scale=1.0, M=None, gn_sums=None, gn_rows_per_img=0,
              gn_groups=32, gn_eps=1e-6, qn_w=None, kn_w=None, qk_eps=1e-6, ln_stats=None, ln_shift=None, ln_scale=None):
    """K2. a [M,K] bf16, w [N,K] bf16 (row strides may exceed K), out [M,N] f32|bf16 per epilogue.
    EPI_GATE_LNRESID_F32: `resid` is x, the residual base modulate(LN(x)) is rebuilt from ln_stats [M,2] (adaln_layernorm's side
    output) and the ln_shift / ln_scale views into the modulation matrix (same row stride as `gate`); `out` may be `resid`."""
    if not w.is_cuda or w.dtype != torch.bfloat16 or w.stride(-1) != 1:
        raise RuntimeError("dfot_b200: `w` must be a CUDA bf16 matrix with unit inner stride")
    if not a.is_cuda or a.dtype != torch.bfloat16 or a.stride(-1) != 1:
        raise RuntimeError("dfot_b200: `a` must be a CUDA bf16 matrix with unit inner stride")
    if not out.is_cuda or out.stride(-1) != 1:
        raise RuntimeError("dfot_b200: `out` must be a CUDA matrix with unit inner stride")
    want = torch.float32 if epilogue in (EPI_F32, EPI_GATE_RESID_F32, EPI_RESID_F32, EPI_GATE_LNRESID_F32) else torch.bfloat16
    if out.dtype != want:
        raise RuntimeError(f"dfot_b200: epilogue {epilogue} writes {want}, got {out.dtype}")
    M = a.shape[0] if M is None else M
    N, K = w.shape
    if a.shape[1] != K or out.shape[0] < M or out.shape[1] != N:
        raise RuntimeError(f"dfot_b200: gemm shape mismatch a{tuple(a.shape)} w{tuple(w.shape)} out{tuple(out.shape)}")
    e = _abi.GemmEpilogue()
    e.bias = _ptr(bias)
    if bias is not None:
        _need(bias, torch.float32, "bias")
    if resid is not None:
        _need(resid, torch.float32, "resid")
        e.resid, e.ld_resid = resid.data_ptr(), resid.stride(0)
    if gate is not None:
        if gate.dtype != torch.float32 or not gate.is_cuda:
            raise RuntimeError("dfot_b200: `gate` must be CUDA f32")
        e.gate, e.ld_gate = gate.data_ptr(), ld_gate
    e.tokens_per_frame = tokens_per_frame
    if epilogue == EPI_GATE_LNRESID_F32:
        _need(ln_stats, torch.float32, "ln_stats")
        for t, name in ((ln_shift, "ln_shift"), (ln_scale, "ln_scale")):
            if t is None or t.dtype != torch.float32 or not t.is_cuda or t.stride(0) != gate.stride(0):
                raise RuntimeError(f"dfot_b200: `{name}` must be a CUDA f32 view with the row stride of `gate`")
        if ln_stats.numel() != 2 * (a.shape[0] if M is None else M):
            raise RuntimeError("dfot_b200: `ln_stats` must be [M, 2] f32")
        e.ln_stats, e.ln_shift, e.ln_scale = ln_stats.data_ptr(), ln_shift.data_ptr(), ln_scale.data_ptr()
    if rope_cs is not None:
        _need(rope_cs, torch.float32, "rope_cs")
        e.rope_cs = rope_cs.data_ptr()
    e.tokens_per_sample, e.model_dim, e.head_dim, e.q_scale = tokens_per_sample, model_dim, head_dim, q_scale
    if qn_w is not None:
        _need(qn_w, torch.float32, "qn_w")
        _need(kn_w, torch.float32, "kn_w")
        e.qn_w, e.kn_w, e.qk_eps = qn_w.data_ptr(), kn_w.data_ptr(), qk_eps
    _gn_side_output(e, gn_sums, gn_rows_per_img, gn_groups, gn_eps, M, N)
    rc = _abi.lib().dfot_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(),
                                   out.stride(0), M, N, K, epilogue, ctypes.byref(e), _stream())
    _abi.check(rc, "gemm_bf16")


import os as _os

# token rows up to which the DiT blocks take the split-K path in latency mode (see splitk_factor); the environment variable
# pins it (benchmarking)
SPLITK_MAX_ROWS = int(_os.environ.get("DFOT_DIT_SPLITK_MAX_ROWS", "1280"))
_latency_mode = _os.environ.get("DFOT_LATENCY_MODE", "0") == "1"


def set_latency_mode(on: bool) -> None:
    """Latency mode (include/dfot_b200.h, off by default): small-batch DiT sampling trades the bit-level batch invariance of
    a forward row for latency — split-K block loop, block-per-row AdaLN, single-tile attention items.  Set it before the
    first forward of a model (CUDA graphs captured earlier keep the kernels they were captured with)."""
    global _latency_mode
    _latency_mode = bool(on)
    if torch.cuda.is_available():
        _abi.check(_abi.lib().dfot_set_latency_mode(int(_latency_mode)), "set_latency_mode")


def latency_mode() -> bool:
    return _latency_mode


def splitk_factor(M: int, N: int, K: int, sms: int = 148) -> int:
    """How many CTAs share the k-loop of one output tile of a [M, N] x K GEMM in the latency regime: as many as keep the
    grid within one wave of the 148 SMs, at least four 64-wide k-blocks each, at most 8.  1 = not worth splitting."""
    tiles = -(-M // 128) * -(-N // 64)
    return max(1, min(sms // max(tiles, 1), (K // 64) // 4, 8))


def gemm_bf16_splitk(a, w, parts, splits, M=None):
    """K2 split-K: a [M,K] bf16, w [N,K] bf16 -> parts [M, splits*N] f32 partial sums (no bias); N % 64 == 0."""
    for t, name in ((a, "a"), (w, "w")):
        if not t.is_cuda or t.dtype != torch.bfloat16 or t.stride(-1) != 1:
            raise RuntimeError(f"dfot_b200: `{name}` must be a CUDA bf16 matrix with unit inner stride")
    _need(parts, torch.float32, "parts")
    M = a.shape[0] if M is None else M
    N, K = w.shape
    if a.shape[1] != K or parts.numel() < M * splits * N:
        raise RuntimeError(f"dfot_b200: split-K shape mismatch a{tuple(a.shape)} w{tuple(w.shape)} parts{tuple(parts.shape)}")
    rc = _abi.lib().dfot_gemm_bf16_splitk(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), parts.data_ptr(), M, N, K,
                                          splits, _stream())
    _abi.check(rc, "gemm_bf16_splitk")


def splitk_gate_resid_adaln(parts, splits, bias, resid, mod, gate_col, shift_col, scale_col, tokens_per_frame, x_out=None,
                            y_f32=None, y_bf16=None, eps=1e-6):
    """x = resid + mod[f, gate_col:] * (sum of the split-K partial sums + bias), then the next AdaLN on x (shift_col < 0:
    none).  resid [M,D] f32, parts [M, splits*D] f32, mod [frames, ld] f32."""
    _need(parts, torch.float32, "parts")
    _need(resid, torch.float32, "resid")
    _need(mod, torch.float32, "mod")
    M, D = resid.shape
    if parts.numel() < M * splits * D:
        raise RuntimeError("dfot_b200: `parts` is smaller than M * splits * D")
    for t, dt, name in ((bias, torch.float32, "bias"), (x_out, torch.float32, "x_out"), (y_f32, torch.float32, "y_f32"),
                        (y_bf16, torch.bfloat16, "y_bf16")):
        if t is not None:
            _need(t, dt, name)
            if name != "bias" and (t.data_ptr() == resid.data_ptr() or tuple(t.shape) != (M, D)):
                raise RuntimeError(f"dfot_b200: `{name}` must be a [M, D] tensor that does not alias `resid`")
    rc = _abi.lib().dfot_splitk_gate_resid_adaln(parts.data_ptr(), splits, _ptr(bias), resid.data_ptr(), mod.data_ptr(),
                                                 mod.shape[-1], gate_col, shift_col, scale_col, _ptr(x_out), _ptr(y_f32),
                                                 _ptr(y_bf16), M, D, tokens_per_frame, eps, _stream())
    _abi.check(rc, "splitk_gate_resid_adaln")


def attention(qkv, out, R, Ntok, heads, head_dim, score_bound=0.0):
    """K3. qkv [R*Ntok, 3*D] bf16 (q rotated+scaled, k rotated), out [R*Ntok, D] bf16 (row stride may exceed D).
    score_bound > 0: an upper bound of |q.k| (pre-scaled, log2 units), e.g. from QK-normalisation — lets the kernel drop
    the running maximum (see include/dfot_b200.h)."""
    _need(qkv, torch.bfloat16, "qkv")
    if not out.is_cuda or out.dtype != torch.bfloat16 or out.stride(-1) != 1 or out.shape[-1] != heads * head_dim:
        raise RuntimeError("dfot_b200: `out` must be a CUDA bf16 [tokens, heads*head_dim] matrix with unit inner stride")
    rc = _abi.lib().dfot_attention_bounded(qkv.data_ptr(), out.data_ptr(), out.stride(0), R, Ntok, heads, head_dim,
                                           float(score_bound), _stream())
    _abi.check(rc, "attention")


def noise_features(levels, out, fourier_freqs=None, fourier_phases=None):
    _need(levels, None, "levels")
    _need(out, torch.bfloat16, "out")
    if levels.dtype not in (torch.int64, torch.float32):
        raise RuntimeError("dfot_b200: noise levels must be int64 or float32")
    n, dim = levels.numel(), out.shape[-1]
    rc = _abi.lib().dfot_noise_features(levels.data_ptr(), _DTYPE_TAG[levels.dtype], _ptr(fourier_freqs),
                                        _ptr(fourier_phases), out.data_ptr(), n, dim, _stream())
    _abi.check(rc, "noise_features")


def silu_sum_bf16(a, b, row_mask, rows_per_mask, out):
    _need(a, torch.float32, "a")
    _need(out, torch.bfloat16, "out")
    if b is not None:
        _need(b, torch.float32, "b")
    if row_mask is not None:
        _need(row_mask, torch.uint8, "row_mask")
    n_rows, D = a.shape
    rc = _abi.lib().dfot_silu_sum_bf16(a.data_ptr(), _ptr(b), _ptr(row_mask), rows_per_mask, out.data_ptr(), n_rows,
                                       D, _stream())
    _abi.check(rc, "silu_sum_bf16")


def patchify_bf16(x, out, frames, C, H, W, p):
    _need(x, None, "x")
    _need(out, torch.bfloat16, "out")
    rc = _abi.lib().dfot_patchify_bf16(x.data_ptr(), _DTYPE_TAG[x.dtype], out.data_ptr(), out.stride(0), frames, C, H,
                                       W, p, _stream())
    _abi.check(rc, "patchify_bf16")


def unpatchify(tok, x, frames, C, H, W, p):
    _need(tok, torch.float32, "tok")
    _need(x, None, "x")
    rc = _abi.lib().dfot_unpatchify(tok.data_ptr(), tok.stride(0), x.data_ptr(), _DTYPE_TAG[x.dtype], frames, C, H, W,
                                    p, _stream())
    _abi.check(rc, "unpatchify")


def cast_bf16(src, out=None):
    _need(src, torch.float32, "src")
    if out is None:
        out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    _need(out, torch.bfloat16, "out")
    rc = _abi.lib().dfot_cast_bf16(src.data_ptr(), out.data_ptr(), src.numel(), _stream())
    _abi.check(rc, "cast_bf16")
    return out


# ------------------------------------------------------------------ U-ViT3DPose kernels
def patch_mix_bf16(y, u, out, R, L, P, Mc):
    """Matrix attention, `qkv_u` factor: y [R*L*P, D] f32, u [P, Mc] f32 -> out [R*Mc*L, D] bf16 (rows (r, c, l))."""
    _need(y, torch.float32, "y")
    _need(u, torch.float32, "u")
    _need(out, torch.bfloat16, "out")
    D = y.shape[-1]
    if y.numel() != R * L * P * D or u.numel() != P * Mc or out.numel() != R * Mc * L * D:
        raise RuntimeError(f"dfot_b200: patch_mix shape mismatch y{tuple(y.shape)} u{tuple(u.shape)} out{tuple(out.shape)}")
    rc = _abi.lib().dfot_patch_mix_bf16(y.data_ptr(), u.data_ptr(), out.data_ptr(), R, L, P, Mc, D, _stream())
    _abi.check(rc, "patch_mix_bf16")


def patch_expand_gate_resid(x, y, z, pu, pb, gate, ld_gate, R, L, P, Mc):
    """Matrix attention, `proj_u` factor + gate + residual: x = y + gate[frame] * (pu^T z + pb); x [R*L*P, D] f32,
    z [R*Mc*L, D] f32, pu [Mc, P] f32, pb [P, D] f32 or None, gate a view into the per-frame modulation matrix;
    y None: no residual, gate None: gate 1."""
    for t, name in ((x, "x"), (z, "z"), (pu, "pu")) + (((y, "y"),) if y is not None else ()):
        _need(t, torch.float32, name)
    if pb is not None:
        _need(pb, torch.float32, "pb")
    if gate is not None and (gate.dtype != torch.float32 or not gate.is_cuda):
        raise RuntimeError("dfot_b200: `gate` must be CUDA f32")
    D = x.shape[-1]
    if (x.numel() != R * L * P * D or (y is not None and y.numel() != x.numel()) or z.numel() != R * Mc * L * D
            or pu.numel() != Mc * P or (pb is not None and pb.numel() != P * D)):
        raise RuntimeError("dfot_b200: patch_expand_gate_resid shape mismatch")
    rc = _abi.lib().dfot_patch_expand_gate_resid(x.data_ptr(), _ptr(y), z.data_ptr(), pu.data_ptr(), _ptr(pb), _ptr(gate),
                                                 ld_gate, R, L, P, Mc, D, _stream())
    _abi.check(rc, "patch_expand_gate_resid")


def conv3x3_bf16(x, w, out, epilogue, bias=None, resid=None, gn_sums=None, gn_groups=32, gn_eps=1e-6):
    """3x3 conv, stride 1, padding 1, implicit GEMM.  x [n,H,W,Cin] bf16 channel-last, w [Cout,3,3,Cin] bf16,
    out [n*H*W, Cout] f32|bf16 per epilogue (EPI_F32, EPI_BF16, EPI_SILU_BF16, EPI_RESID_F32)."""
    _need(x, torch.bfloat16, "x")
    _need(w, torch.bfloat16, "w")
    _need(out, None, "out")
    n, H, W, Cin = x.shape
    Cout = w.shape[0]
    if tuple(w.shape) != (Cout, 3, 3, Cin) or out.numel() != n * H * W * Cout:
        raise RuntimeError(f"dfot_b200: conv3x3 shape mismatch x{tuple(x.shape)} w{tuple(w.shape)} out{tuple(out.shape)}")
    want = torch.float32 if epilogue in (EPI_F32, EPI_RESID_F32) else torch.bfloat16
    if out.dtype != want:
        raise RuntimeError(f"dfot_b200: epilogue {epilogue} writes {want}, got {out.dtype}")
    e = _abi.GemmEpilogue()
    if bias is not None:
        _need(bias, torch.float32, "bias")
        e.bias = bias.data_ptr()
    if resid is not None:
        _need(resid, torch.float32, "resid")
        e.resid, e.ld_resid = resid.data_ptr(), Cout
    e.tokens_per_frame = 1
    _gn_side_output(e, gn_sums, H * W, gn_groups, gn_eps, n * H * W, Cout)
    rc = _abi.lib().dfot_conv3x3_bf16(x.data_ptr(), w.data_ptr(), out.data_ptr(), Cout, n, H, W, Cin, Cout, epilogue,
                                      ctypes.byref(e), _stream())
    _abi.check(rc, "conv3x3_bf16")


def conv3d_causal_bf16(x, w, out, epilogue, bias=None, resid=None):
    """Causal kt x 3 x 3 conv along a frame axis (the reference VideoVAE's PaddedConv3D).  x [n_out + kt - 1, H, W, Cin]
    bf16 channel-last with frames j .. j+kt-1 = output frame j's causal window, w [Cout, kt, 3, 3, Cin] bf16,
    out [n_out*H*W, Cout] f32|bf16 per epilogue (EPI_F32, EPI_BF16, EPI_SILU_BF16, EPI_RESID_F32)."""
    _need(x, torch.bfloat16, "x")
    _need(w, torch.bfloat16, "w")
    _need(out, None, "out")
    n_in, H, W, Cin = x.shape
    Cout, kt = w.shape[0], w.shape[1]
    n_out = n_in - kt + 1
    if tuple(w.shape) != (Cout, kt, 3, 3, Cin) or n_out < 1 or out.numel() != n_out * H * W * Cout:
        raise RuntimeError(f"dfot_b200: conv3d shape mismatch x{tuple(x.shape)} w{tuple(w.shape)} out{tuple(out.shape)}")
    want = torch.float32 if epilogue in (EPI_F32, EPI_RESID_F32) else torch.bfloat16
    if out.dtype != want:
        raise RuntimeError(f"dfot_b200: epilogue {epilogue} writes {want}, got {out.dtype}")
    e = _abi.GemmEpilogue()
    if bias is not None:
        _need(bias, torch.float32, "bias")
        e.bias = bias.data_ptr()
    if resid is not None:
        _need(resid, torch.float32, "resid")
        e.resid, e.ld_resid = resid.data_ptr(), Cout
    e.tokens_per_frame = 1
    rc = _abi.lib().dfot_conv3d_causal_bf16(x.data_ptr(), w.data_ptr(), out.data_ptr(), Cout, n_out, H, W, Cin, Cout, kt,
                                            epilogue, ctypes.byref(e), _stream())
    _abi.check(rc, "conv3d_causal_bf16")


# ---------------------------------------------------------------- VAE-decode building blocks (clips [B, 2 + T, H, W, C])
def groupnorm_stats_strided(x, sums, n_img, HW, img_stride, C, groups=32, eps=1e-6):
    """GroupNorm statistics over strided images: image i = HW rows of C channels starting at x.flatten()[i*img_stride]."""
    _need(x, None, "x")
    _need(sums, torch.float64, "sums")
    if sums.numel() != n_img * groups * 3:
        raise RuntimeError("dfot_b200: groupnorm_stats_strided workspace must hold n_img*groups*3 doubles")
    rc = _abi.lib().dfot_groupnorm_stats_strided(x.data_ptr(), _DTYPE_TAG[x.dtype], sums.data_ptr(), n_img, HW, img_stride,
                                                 C, groups, eps, _stream())
    _abi.check(rc, "groupnorm_stats_strided")


def groupnorm_apply_bf16(x, sums, gamma, beta, out, n_img, HW, img_stride, C, groups=32, silu=True):
    """y = GroupNorm(x) * gamma + beta (then x*sigmoid(x) if silu) -> bf16 at the same strided offsets of `out`."""
    _need(x, torch.float32, "x")
    _need(out, torch.bfloat16, "out")
    _need(gamma, torch.float32, "gamma")
    _need(beta, torch.float32, "beta")
    rc = _abi.lib().dfot_groupnorm_apply_bf16(x.data_ptr(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                              out.data_ptr(), n_img, HW, img_stride, C, groups, 1 if silu else 0, _stream())
    _abi.check(rc, "groupnorm_apply_bf16")


def vae_upsample2x_bf16(x, out, B, T_in, H, W, C, temporal):
    """fp32 clip [B, 2+T_in, H, W, C] -> bf16 clip [B, 2+T_out, 2H, 2W, C]; temporal: T_out = 2*T_in - 1 (first frame
    bilinear, the others trilinear), else nearest with T_out = T_in; pad slots of `out` are filled."""
    _need(x, torch.float32, "x")
    _need(out, torch.bfloat16, "out")
    T_out = 2 * T_in - 1 if temporal else T_in
    if x.numel() != B * (2 + T_in) * H * W * C or out.numel() != B * (2 + T_out) * 4 * H * W * C:
        raise RuntimeError("dfot_b200: vae_upsample2x shape mismatch")
    rc = _abi.lib().dfot_vae_upsample2x_bf16(x.data_ptr(), out.data_ptr(), B, T_in, H, W, C, 1 if temporal else 0, _stream())
    _abi.check(rc, "vae_upsample2x_bf16")


def upsample2x_nearest_bf16(x, out, n_img, H, W, C):
    """fp32 images [n_img, H, W, C] -> bf16 [n_img, 2H, 2W, C], nearest x2 (the ImageVAE decoder's Upsample)."""
    _need(x, torch.float32, "x")
    _need(out, torch.bfloat16, "out")
    if x.numel() != n_img * H * W * C or out.numel() != 4 * x.numel():
        raise RuntimeError("dfot_b200: upsample2x_nearest shape mismatch")
    rc = _abi.lib().dfot_upsample2x_nearest_bf16(x.data_ptr(), out.data_ptr(), n_img, H, W, C, _stream())
    _abi.check(rc, "upsample2x_nearest_bf16")


def vae_fill_pad_frames(x, B, T, frame_elems):
    """pad slots (2 per clip) of a bf16 clip [B, 2+T, ...] <- the clip's first frame."""
    _need(x, torch.bfloat16, "x")
    if x.numel() != B * (2 + T) * frame_elems:
        raise RuntimeError("dfot_b200: vae_fill_pad_frames shape mismatch")
    rc = _abi.lib().dfot_vae_fill_pad_frames(x.data_ptr(), B, T, frame_elems, _stream())
    _abi.check(rc, "vae_fill_pad_frames")


def softmax_rows_bf16(s, p, scale=1.0):
    """p = softmax(scale * s, dim=-1): s [rows, n] f32 (row stride s.stride(0)) -> p [rows, n] bf16."""
    if s.dtype != torch.float32 or p.dtype != torch.bfloat16 or not s.is_cuda or s.shape != p.shape or s.stride(1) != 1 or p.stride(1) != 1:
        raise RuntimeError("dfot_b200: softmax_rows_bf16 needs CUDA f32 logits and a bf16 output of the same shape")
    rc = _abi.lib().dfot_softmax_rows_bf16(s.data_ptr(), s.stride(0), p.data_ptr(), p.stride(0), s.shape[0], s.shape[1],
                                           float(scale), _stream())
    _abi.check(rc, "softmax_rows_bf16")


def groupnorm_stats(x, sums, n_img, HW, C, groups=32, eps=1e-6):
    """x [n_img*HW, C] f32|bf16 channel-last → sums [n_img, groups, 3] f64 workspace: (sum, sum of squares) pairs
    followed by the finalised f32 (mean, rstd) pairs that groupnorm_silu_bf16 reads."""
    _need(x, None, "x")
    _need(sums, torch.float64, "sums")
    if sums.numel() != n_img * groups * 3 or x.numel() != n_img * HW * C:
        raise RuntimeError("dfot_b200: groupnorm_stats size mismatch (workspace = 3 * n_img * groups doubles)")
    rc = _abi.lib().dfot_groupnorm_stats(x.data_ptr(), _DTYPE_TAG[x.dtype], sums.data_ptr(), n_img, HW, C, groups, eps,
                                         _stream())
    _abi.check(rc, "groupnorm_stats")


def groupnorm_silu_bf16(x, sums, gamma, beta, out, n_img, HW, C, groups=32, mod_img=None, scale_col=0,
                        shift_col=0, mod_pix=None, img_map=None):
    _need(x, None, "x")
    _need(sums, torch.float64, "sums")
    _need(gamma, torch.float32, "gamma")
    _need(beta, torch.float32, "beta")
    _need(out, torch.bfloat16, "out")
    if mod_img is not None:
        _need(mod_img, torch.float32, "mod_img")
    if mod_pix is not None:
        _need(mod_pix, torch.bfloat16, "mod_pix")
        _need(img_map, torch.int32, "img_map")
    rc = _abi.lib().dfot_groupnorm_silu_bf16(
        x.data_ptr(), _DTYPE_TAG[x.dtype], sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(mod_img),
        0 if mod_img is None else mod_img.shape[-1], scale_col, shift_col, _ptr(mod_pix), _ptr(img_map),
        out.data_ptr(), n_img, HW, C, groups, _stream())
    _abi.check(rc, "groupnorm_silu_bf16")


def rmsnorm_film_bf16(x, weight, mod_img, scale_col, shift_col, tokens_per_img, out, mod_pix=None, img_map=None,
                      eps=1e-6):
    _need(x, torch.float32, "x")
    _need(weight, torch.float32, "weight")
    _need(mod_img, torch.float32, "mod_img")
    _need(out, torch.bfloat16, "out")
    if mod_pix is not None:
        _need(mod_pix, torch.bfloat16, "mod_pix")
        _need(img_map, torch.int32, "img_map")
    M, D = x.shape
    rc = _abi.lib().dfot_rmsnorm_film_bf16(x.data_ptr(), weight.data_ptr(), eps, mod_img.data_ptr(), mod_img.shape[-1],
                                           scale_col, shift_col, _ptr(mod_pix), _ptr(img_map), out.data_ptr(), M, D,
                                           tokens_per_img, _stream())
    _abi.check(rc, "rmsnorm_film_bf16")


def qk_norm_rope(qkv, q_weight, k_weight, rope_cs, tokens_per_sample, heads, head_dim, q_scale, eps=1e-6):
    """In place on qkv [M, ld] bf16 (columns [q | k | v ...]); row stride may exceed 3*heads*head_dim."""
    if not qkv.is_cuda or qkv.dtype != torch.bfloat16 or qkv.stride(-1) != 1:
        raise RuntimeError("dfot_b200: `qkv` must be a CUDA bf16 matrix with unit inner stride")
    _need(q_weight, torch.float32, "q_weight")
    _need(k_weight, torch.float32, "k_weight")
    _need(rope_cs, torch.float32, "rope_cs")
    rc = _abi.lib().dfot_qk_norm_rope(qkv.data_ptr(), qkv.stride(0), q_weight.data_ptr(), k_weight.data_ptr(), eps,
                                      rope_cs.data_ptr(), tokens_per_sample, qkv.shape[0], heads, head_dim, q_scale,
                                      _stream())
    _abi.check(rc, "qk_norm_rope")


def avgpool2x2(x, out, n_img, H, W, C):
    _need(x, None, "x")
    _need(out, None, "out")
    rc = _abi.lib().dfot_avgpool2x2(x.data_ptr(), _DTYPE_TAG[x.dtype], out.data_ptr(), _DTYPE_TAG[out.dtype], n_img, H,
                                    W, C, _stream())
    _abi.check(rc, "avgpool2x2")


def sub_bf16(a, b, out):
    _need(a, torch.float32, "a")
    _need(b, torch.float32, "b")
    _need(out, torch.bfloat16, "out")
    rc = _abi.lib().dfot_sub_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream())
    _abi.check(rc, "sub_bf16")


def upsample2x_add(low, skip, out, n_img, H, W, C):
    """out [n,H,W,C] = nearest2x(low [n,H/2,W/2,C]) + skip [n,H,W,C] (all f32, channel-last)."""
    _need(low, torch.float32, "low")
    _need(skip, torch.float32, "skip")
    _need(out, torch.float32, "out")
    rc = _abi.lib().dfot_upsample2x_add(low.data_ptr(), skip.data_ptr(), out.data_ptr(), n_img, H, W, C, _stream())
    _abi.check(rc, "upsample2x_add")


def pose_ray_patches(cams, freq_scale, out, frames, res, p):
    _need(cams, torch.float32, "cams")
    _need(freq_scale, torch.float32, "freq_scale")
    if not out.is_cuda or out.dtype != torch.bfloat16 or out.dim() != 2 or out.stride(1) != 1:
        raise RuntimeError("dfot_b200: `out` must be a CUDA bf16 matrix with unit inner stride (rows may be padded)")
    if out.shape[0] != frames * (res // p) ** 2 or out.shape[1] != p * p * 12 * freq_scale.numel():
        raise RuntimeError(f"dfot_b200: pose_ray_patches `out` has shape {tuple(out.shape)}")
    rc = _abi.lib().dfot_pose_ray_patches(cams.data_ptr(), freq_scale.data_ptr(), freq_scale.numel(), out.data_ptr(),
                                          out.stride(0), frames, res, p, _stream())
    _abi.check(rc, "pose_ray_patches")


# ---------------------------------------------------------------- DC-AE decoder glue (include/dfot_b200.h)
def relu_bf16(x):
    _need(x, torch.bfloat16, "x")
    rc = _abi.lib().dfot_relu_bf16(x.data_ptr(), x.numel(), _stream())
    _abi.check(rc, "relu_bf16")


def pixel_shuffle2x(conv, C, n_img, H, W, shortcut=None, repeats=1, out_f32=None, out_bf16=None):
    """conv [n*H*W, >= 4C] f32 (+ shortcut [n*H*W, 4C / repeats] f32) -> [n, 2H, 2W, C] f32 and / or bf16."""
    if not conv.is_cuda or conv.dtype != torch.float32 or conv.stride(-1) != 1 or conv.shape[0] != n_img * H * W:
        raise RuntimeError("dfot_b200: `conv` must be a CUDA f32 [n*H*W, >=4C] matrix with unit inner stride")
    Cx = 0
    if shortcut is not None:
        _need(shortcut, torch.float32, "shortcut")
        Cx = shortcut.shape[-1]
    for t, dt, n in ((out_f32, torch.float32, "out_f32"), (out_bf16, torch.bfloat16, "out_bf16")):
        if t is not None:
            _need(t, dt, n)
            if t.numel() != n_img * 4 * H * W * C:
                raise RuntimeError(f"dfot_b200: `{n}` must hold n*2H*2W*C elements")
    rc = _abi.lib().dfot_pixel_shuffle2x(conv.data_ptr(), conv.stride(0), _ptr(shortcut), Cx, repeats, _ptr(out_f32),
                                         _ptr(out_bf16), n_img, H, W, C, _stream())
    _abi.check(rc, "pixel_shuffle2x")


def linear_attention_relu(qkv, out, n_img, HW, heads, head_dim, eps=1e-15):
    for t, n in ((qkv, "qkv"), (out, "out")):
        if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.stride(-1) != 1:
            raise RuntimeError(f"dfot_b200: `{n}` must be a CUDA f32 matrix with unit inner stride")
    if qkv.shape[0] != n_img * HW or out.shape[0] != n_img * HW:
        raise RuntimeError("dfot_b200: linear_attention_relu expects [n_img*HW, ...] matrices")
    rc = _abi.lib().dfot_linear_attention_relu(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), n_img, HW, heads,
                                               head_dim, float(eps), _stream())
    _abi.check(rc, "linear_attention_relu")


def dwconv3x3_glu_bf16(x, w, b, out, n_img, H, W, Ch):
    _need(x, torch.bfloat16, "x")
    _need(w, torch.float32, "w")
    _need(b, torch.float32, "b")
    _need(out, torch.bfloat16, "out")
    if x.numel() != n_img * H * W * 2 * Ch or out.numel() != n_img * H * W * Ch or w.numel() != 18 * Ch or b.numel() != 2 * Ch:
        raise RuntimeError("dfot_b200: dwconv3x3_glu shape mismatch")
    rc = _abi.lib().dfot_dwconv3x3_glu_bf16(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n_img, H, W, Ch, _stream())
    _abi.check(rc, "dwconv3x3_glu_bf16")


def rmsnorm_affine(x, w, b, eps, resid=None, relu=False, out_f32=None, out_bf16=None):
    if not x.is_cuda or x.dtype != torch.float32 or x.stride(-1) != 1:
        raise RuntimeError("dfot_b200: `x` must be a CUDA f32 matrix with unit inner stride")
    _need(w, torch.float32, "w")
    _need(b, torch.float32, "b")
    M, C = x.shape
    if resid is not None:
        _need(resid, torch.float32, "resid")
    if out_f32 is not None:
        _need(out_f32, torch.float32, "out_f32")
    if out_bf16 is not None:
        _need(out_bf16, torch.bfloat16, "out_bf16")
    rc = _abi.lib().dfot_rmsnorm_affine(x.data_ptr(), x.stride(0), w.data_ptr(), b.data_ptr(), float(eps), _ptr(resid),
                                        1 if relu else 0, _ptr(out_f32), _ptr(out_bf16), M, C, _stream())
    _abi.check(rc, "rmsnorm_affine")
