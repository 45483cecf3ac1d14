"""ORACLE (test infrastructure): noise schedules and scheduling matrices.

Restates
  algorithms/dfot/diffusion/noise_schedule.py:6-159        (beta schedules)
  algorithms/dfot/diffusion/discrete_diffusion.py:94-168   (_build_buffer)
  algorithms/dfot/diffusion/discrete_diffusion.py:379-384  (ddim_idx_to_noise_level)
  algorithms/common/base_pytorch_video_algo.py:877-947     (scheduling matrices)
"""
import math
from typing import Dict

import numpy as np
import torch


# --------------------------------------------------------------------------
# alphas_cumprod families (float64), noise_schedule.py:38-140
# --------------------------------------------------------------------------
def _ac_cosine(timesteps: int, s: float = 0.008) -> torch.Tensor:
    # noise_schedule.py:38-47
    t = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps
    ac = torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** 2
    return (ac / ac[0])[1:]


def _ac_cosine_simple_diffusion(timesteps: int, logsnr_min: float = -15.0, logsnr_max: float = 15.0,
                                shifted: float = 1.0, interpolated: bool = False) -> torch.Tensor:
    # noise_schedule.py:50-84
    t_min = torch.atan(torch.exp(-0.5 * torch.tensor(logsnr_max, dtype=torch.float64)))
    t_max = torch.atan(torch.exp(-0.5 * torch.tensor(logsnr_min, dtype=torch.float64)))
    t = torch.linspace(0, 1, timesteps, dtype=torch.float64)
    logsnr = -2 * torch.log(torch.tan(t_min + t * (t_max - t_min)))
    if shifted != 1.0:
        moved = logsnr + 2 * torch.log(torch.tensor(shifted, dtype=torch.float64))
        logsnr = t * logsnr + (1 - t) * moved if interpolated else moved
    return 1 / (1 + torch.exp(-logsnr))


def _ac_linear_alphas(timesteps: int) -> torch.Tensor:
    # noise_schedule.py:87-93
    t = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps
    return (1 - t)[1:]


def _ac_beta_linear(timesteps: int, start: float = 0.0001, end: float = 0.02) -> torch.Tensor:
    # noise_schedule.py:96-104
    return (1 - torch.linspace(start, end, timesteps, dtype=torch.float64)).cumprod(dim=0)


def _ac_sigmoid(timesteps: int, start=-3, end=3, tau=1) -> torch.Tensor:
    # noise_schedule.py:107-122
    t = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps
    v_start = torch.tensor(start / tau).sigmoid()
    v_end = torch.tensor(end / tau).sigmoid()
    ac = (-((t * (end - start) + start) / tau).sigmoid() + v_end) / (v_end - v_start)
    return (ac / ac[0])[1:]


def _ac_sd(timesteps: int, start: float = 0.00085, end: float = 0.0120) -> torch.Tensor:
    # noise_schedule.py:125-132
    betas = torch.linspace(start ** 0.5, end ** 0.5, timesteps, dtype=torch.float64) ** 2
    return (1 - betas).cumprod(dim=0)


def _zero_terminal_snr(ac: torch.Tensor) -> torch.Tensor:
    # noise_schedule.py:145-159
    r = torch.sqrt(ac)
    r0, rT = r[0].clone(), r[-1].clone()
    r = r - rT
    r = r * (r0 / r[0])
    ac = r ** 2
    assert ac[-1] == 0
    return ac


_FAMILIES = {
    "alphas_cumprod_linear": _ac_linear_alphas,
    "cosine": _ac_cosine,
    "cosine_simple_diffusion": _ac_cosine_simple_diffusion,
    "sigmoid": _ac_sigmoid,
    "sd": _ac_sd,
    "linear": _ac_beta_linear,
}


def make_betas(schedule: str, timesteps: int, zero_terminal_snr: bool, shift: float = 1.0,
               clip_min: float = 1e-9, **kwargs) -> torch.Tensor:
    """noise_schedule.py:6-35 (float64 betas)."""
    ac = _FAMILIES[schedule](timesteps=timesteps, **kwargs)
    cosine_like = schedule in ("cosine", "cosine_simple_diffusion")
    if zero_terminal_snr and not cosine_like:
        ac = _zero_terminal_snr(ac)
    if shift != 1.0 and schedule != "cosine_simple_diffusion":
        s2 = shift ** 2
        ac = (s2 * ac) / (s2 * ac + 1 - ac)  # noise_schedule.py:135-142
    alphas = torch.cat([ac[0:1], ac[1:] / ac[:-1]])
    return torch.clip(1 - alphas, clip_min, 1.0)


def diffusion_buffers(diff_cfg: dict) -> Dict[str, torch.Tensor]:
    """fp32 tables used at sampling time (discrete_diffusion.py:94-168).

    diff_cfg: the ``algorithm.diffusion`` config sub-tree (plain dict)."""
    kwargs = dict(diff_cfg.get("schedule_fn_kwargs") or {})
    betas = make_betas(diff_cfg["beta_schedule"], diff_cfg["timesteps"],
                       zero_terminal_snr=diff_cfg["objective"] != "pred_noise", **kwargs)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = torch.cat([torch.ones(1, dtype=ac.dtype), ac[:-1]])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)          # q(x_{t-1} | x_t, x_0), discrete_diffusion.py:135-157
    out = {
        "posterior_log_variance_clipped": torch.log(post_var.clamp(min=1e-20)),
        "posterior_mean_coef1": betas * torch.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac),
        "betas": betas,
        "alphas_cumprod": ac,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / ac - 1),
    }
    if diff_cfg["loss_weighting"]["strategy"] == "sigmoid":
        out["logsnr"] = torch.log(ac / (1 - ac))
    return {k: v.to(torch.float32) for k, v in out.items()}


def ddim_idx_to_noise_level(indices: torch.Tensor, timesteps: int, sampling_timesteps: int) -> torch.Tensor:
    """discrete_diffusion.py:379-384 — fp32 linspace then truncation."""
    real_steps = torch.linspace(-1, timesteps - 1, sampling_timesteps + 1).long()
    return real_steps[indices.flatten()].view(indices.shape)


# --------------------------------------------------------------------------
# scheduling matrices, base_pytorch_video_algo.py:877-947
# --------------------------------------------------------------------------
def pyramid_index_matrix(horizon: int, steps: int, uncertainty_scale: float = 1.0) -> np.ndarray:
    # base_pytorch_video_algo.py:940-947
    height = steps + int((horizon - 1) * uncertainty_scale) + 1
    m = np.arange(height, dtype=np.int64)[:, None]
    t = np.array([int(i * uncertainty_scale) for i in range(horizon)], dtype=np.int64)[None, :]
    return np.clip(steps + t - m, 0, steps)


def interleaved_index_matrix(horizon: int, interleaved_size: int, steps: int) -> np.ndarray:
    # base_pytorch_video_algo.py:915-938
    cols = []
    full_len = steps + interleaved_size
    for i in range(horizon):
        start = i % interleaved_size + 1
        col = [steps] * start
        for j in range(steps):
            idx = max(steps - start - interleaved_size * j, 0)
            if idx == 0:
                col += [0] * (full_len - len(col))
                break
            col += [idx] * interleaved_size
        cols.append(col)
    return np.array(cols).T


def scheduling_matrix(kind: str, horizon: int, padding: int, timesteps: int, sampling_timesteps: int) -> torch.Tensor:
    """base_pytorch_video_algo.py:877-913 → int64 [M, horizon+padding] of noise levels."""
    S = sampling_timesteps
    if kind in ("full_sequence", "gibbs"):
        idx = np.arange(S, -1, -1)[:, None].repeat(horizon, axis=1)
    elif kind == "autoregressive":
        idx = pyramid_index_matrix(horizon, S)
    elif kind == "interleaved":
        idx = interleaved_index_matrix(horizon, 3, S)  # size hard-coded at :892-894
    else:
        raise ValueError(kind)
    lv = ddim_idx_to_noise_level(torch.from_numpy(idx).long(), timesteps, S)
    if kind == "gibbs":  # :901-906
        n = lv.shape[0]
        lv = lv.repeat_interleave(horizon, dim=0)
        for i in range(1, n):
            for j in range(horizon):
                lv[i * horizon + j, j + 1:] = lv[(i - 1) * horizon + horizon - 1, j + 1:]
    return torch.nn.functional.pad(lv, (0, padding, 0, 0), value=timesteps - 1)


def refine_scheduling_matrix(horizon: int, goback_length: int, n_goback: int, padding: int, timesteps: int,
                             sampling_timesteps: int) -> torch.Tensor:
    """base_pytorch_video_algo.py:949-976 (fork-only refinement sampling, full_sequence only): the DDIM index walk
    S, S-1, ..., 0 where, after every index t in range(1, S - goback_length, goback_length), the walk goes back up to
    t + goback_length and down to t again, n_goback times; indices -> levels; pad columns = timesteps - 1."""
    S = sampling_timesteps
    goback = set(range(1, S - goback_length, goback_length))
    walk = []
    for t in range(S, -1, -1):
        walk.append(t)
        if t in goback:
            for _ in range(n_goback):
                walk += list(range(t + 1, t + goback_length + 1)) + list(range(t + goback_length - 1, t - 1, -1))
    lv = ddim_idx_to_noise_level(torch.tensor(walk).long(), timesteps, S)[:, None].repeat(1, horizon)
    return torch.nn.functional.pad(lv, (0, padding, 0, 0), value=timesteps - 1)
