"""Host-side planning of a denoising window for the fused K4 kernel.

Everything integer / per-frame-scalar about a window is known before its first step: the scheduling
matrix (base_pytorch_video_algo.py:877-947), the evolution of the context mask (dfot_video.py:675-679),
the history-guidance branch table of every step (history_guidance.py:357-437) and hence the DDIM
coefficients (discrete_diffusion.py:464-478) and the prepare instructions (history_guidance.py:446-543,
929-973) of every (step, branch row, frame).  The reference re-derives them on the device every step with
~40 tiny ATen kernels and several host syncs; here they are computed once per window in numpy/torch on
the host, packed into the C structs of include/dfot_b200.h and uploaded with one copy per table.
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

UPDATE_DTYPE = np.dtype([("a", "<f4"), ("b", "<f4"), ("sigma", "<f4"), ("w", "<f4"), ("clip", "<f4"),
                         ("generate", "<i4")])
PREPARE_DTYPE = np.dtype([("mode", "<i4"), ("noise_row", "<i4"), ("qa", "<f4"), ("qb", "<f4")])
MODE_COPY, MODE_QSAMPLE, MODE_NOISE = 0, 1, 2


@dataclass
class HostTables:
    """fp32 diffusion tables as numpy (copies of the registered buffers)."""
    alphas_cumprod: np.ndarray
    sqrt_alphas_cumprod: np.ndarray
    sqrt_one_minus_alphas_cumprod: np.ndarray
    sqrt_recip_alphas_cumprod: np.ndarray
    sqrt_recipm1_alphas_cumprod: np.ndarray
    logsnr: Optional[np.ndarray]
    objective: str
    eta: float
    clip_noise: float
    timesteps: int
    is_ddim: bool = True                                   # False: sampling_timesteps == timesteps, DDPM ancestral steps
    posterior_mean_coef1: Optional[np.ndarray] = None
    posterior_mean_coef2: Optional[np.ndarray] = None
    posterior_log_variance_clipped: Optional[np.ndarray] = None

    @property
    def uses_step_noise(self) -> bool:
        """Does the update consume the per-step noise draw?  (It is DRAWN either way, to keep the RNG stream aligned.)"""
        return self.eta != 0 or not self.is_ddim


def step_update_table(tb: "HostTables", frm: np.ndarray, to: np.ndarray, weight: np.ndarray,
                      generate: np.ndarray) -> np.ndarray:
    """Per (row, frame) coefficients of one sampling step: DDIM, or DDPM when sampling_timesteps == timesteps
    (discrete_diffusion.py:386-421 dispatch)."""
    return ddim_update_table(tb, frm, to, weight, generate) if tb.is_ddim else ddpm_update_table(tb, frm, weight, generate)


def ddpm_update_table(tb: "HostTables", frm: np.ndarray, weight: np.ndarray, generate: np.ndarray) -> np.ndarray:
    """DDPM ancestral step (discrete_diffusion.py:423-452, 231-240) in the same  x' = a*x + b*g(out) + sigma*noise  form:
    x0 = ax*x + ao*g(out) by objective, mean = coef1[k]*x0 + coef2[k]*x, sigma = exp(0.5*log_var[k]) where k > 0 (the
    reference zeroes the noise at k == 0), and the frame is kept only where the current level is -1."""
    k = np.clip(frm, 0, None)
    f8 = np.float64
    c1, c2 = tb.posterior_mean_coef1.astype(f8)[k], tb.posterior_mean_coef2.astype(f8)[k]
    sa = tb.sqrt_alphas_cumprod.astype(f8)[k]
    sb = tb.sqrt_one_minus_alphas_cumprod.astype(f8)[k]
    if tb.objective == "pred_v":
        ax, ao, clip = sa, -sb, np.zeros_like(sa)
    elif tb.objective == "pred_x0":
        ax, ao, clip = np.zeros_like(sa), np.ones_like(sa), np.zeros_like(sa)
    elif tb.objective == "pred_noise":
        ax, ao = tb.sqrt_recip_alphas_cumprod.astype(f8)[k], -tb.sqrt_recipm1_alphas_cumprod.astype(f8)[k]
        clip = np.full_like(sa, tb.clip_noise)
    else:
        raise ValueError(f"unknown objective {tb.objective}")
    sigma = np.where(k > 0, np.exp(0.5 * tb.posterior_log_variance_clipped.astype(f8)[k]), 0.0)
    keep = frm == -1
    out = np.zeros(frm.shape, dtype=UPDATE_DTYPE)
    out["a"] = np.where(keep, 1.0, c1 * ax + c2)
    out["b"] = np.where(keep, 0.0, c1 * ao)
    out["sigma"] = np.where(keep, 0.0, sigma)
    out["w"] = weight
    out["clip"] = clip
    out["generate"] = generate
    return out


def ddim_update_table(tb: HostTables, frm: np.ndarray, to: np.ndarray, weight: np.ndarray,
                      generate: np.ndarray) -> np.ndarray:
    """Per (row, frame) coefficients of  x' = a*x + b*g(out) + sigma*noise  (discrete_diffusion.py:454-538).

    The reference forms x0 and eps separately and then x' = sqrt(ᾱ')·x0 + c·eps + σ·n; both are linear in
    (x, out), so the per-frame scalars are folded here (float64, rounded once to fp32)."""
    k = np.clip(frm, 0, None)
    f8 = np.float64
    ac = tb.alphas_cumprod.astype(f8)
    alpha = ac[k]
    to_c = np.clip(to, 0, None)
    alpha_next = np.where(to < 0, 1.0, ac[to_c])
    with np.errstate(divide="ignore", invalid="ignore"):
        sigma = np.where(to < 0, 0.0, tb.eta * np.sqrt((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)))
        sigma = np.nan_to_num(sigma, nan=0.0) if tb.eta == 0 else sigma
        c = np.sqrt(1 - alpha_next - sigma ** 2)
        san = np.sqrt(alpha_next)
        sa = tb.sqrt_alphas_cumprod.astype(f8)[k]
        sb = tb.sqrt_one_minus_alphas_cumprod.astype(f8)[k]
        if tb.objective == "pred_v":
            a, b = san * sa + c * sb, c * sa - san * sb
            clip = np.zeros_like(a)
        elif tb.objective == "pred_x0":
            a, b = c / sb, san - c * sa / sb
            clip = np.zeros_like(a)
        elif tb.objective == "pred_noise":
            sr = tb.sqrt_recip_alphas_cumprod.astype(f8)[k]
            srm1 = tb.sqrt_recipm1_alphas_cumprod.astype(f8)[k]
            a, b = san * sr, c - san * srm1
            clip = np.full_like(a, tb.clip_noise)
        else:
            raise ValueError(f"unknown objective {tb.objective}")
    keep = frm == to  # discrete_diffusion.py:530-536
    out = np.zeros(frm.shape, dtype=UPDATE_DTYPE)
    out["a"] = np.where(keep, 1.0, a)
    out["b"] = np.where(keep, 0.0, b)
    out["sigma"] = np.where(keep, 0.0, sigma)
    out["w"] = weight
    out["clip"] = clip
    out["generate"] = generate
    return out


def model_levels(tb: HostTables, frm: np.ndarray, continuous: bool, precond_scale: float) -> np.ndarray:
    """What the backbone is fed as noise level: k = clamp(from, 0) (discrete_diffusion.py:464) or
    precond_scale * logsnr[k] (continuous_diffusion.py:118-121)."""
    k = np.clip(frm, 0, None)
    if continuous:
        return (np.float32(precond_scale) * tb.logsnr[k]).astype(np.float32)
    return k.astype(np.int64)


@dataclass
class StepPlan:
    nfe: int
    levels: np.ndarray            # [B*nfe, T] int64 | f32 — backbone noise-level input
    cond_mask: Optional[np.ndarray]  # [B*nfe] bool or None
    update: np.ndarray            # [B*nfe, T] UPDATE_DTYPE
    prepare: np.ndarray           # [B*nfe, T] PREPARE_DTYPE (how THIS step's model input is built)
    n_hist_rows: int              # rows of q_sample noise the reference draws (B or B*h), 0 = none
    draws_excluded_noise: bool    # full manager always draws randn_like((B*h, g, T, ...))
    context_mask: np.ndarray      # [B, T] mask in force during this step (after the 0→2 update)
    levels_from: np.ndarray       # [B*nfe, T] int64 (for traces / parity tests)
    levels_to: np.ndarray


def to_device_bytes(arr: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).to(device, non_blocking=True)
