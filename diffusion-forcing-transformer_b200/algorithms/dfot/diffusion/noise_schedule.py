"""Host-side noise schedules (float64 → float32 tables), same menu and semantics as the reference's
algorithms/dfot/diffusion/noise_schedule.py:6-159.  Runs once at construction; nothing here is on the
per-step path."""
import math

import torch

_PI_2 = math.pi * 0.5


def _grid(timesteps: int) -> torch.Tensor:
    return torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps


def alphas_cumprod_for(schedule: str, timesteps: int, **kw) -> torch.Tensor:
    if schedule == "cosine":
        s = kw.get("s", 0.008)
        f = torch.cos((_grid(timesteps) + s) / (1 + s) * _PI_2) ** 2
        return (f / f[0])[1:]
    if schedule == "cosine_simple_diffusion":
        lo, hi = kw.get("logsnr_min", -15.0), kw.get("logsnr_max", 15.0)
        shifted, interpolated = kw.get("shifted", 1.0), kw.get("interpolated", False)
        t_min = torch.atan(torch.exp(-0.5 * torch.tensor(hi, dtype=torch.float64)))
        t_max = torch.atan(torch.exp(-0.5 * torch.tensor(lo, dtype=torch.float64)))
        t = torch.linspace(0, 1, timesteps, dtype=torch.float64)
        logsnr = -2 * torch.log(torch.tan(t_min + t * (t_max - t_min)))
        if shifted != 1.0:
            moved = logsnr + 2 * torch.log(torch.tensor(shifted, dtype=torch.float64))
            logsnr = t * logsnr + (1 - t) * moved if interpolated else moved
        return 1 / (1 + torch.exp(-logsnr))
    if schedule == "alphas_cumprod_linear":
        return (1 - _grid(timesteps))[1:]
    if schedule == "linear":
        betas = torch.linspace(kw.get("start", 0.0001), kw.get("end", 0.02), timesteps, dtype=torch.float64)
        return (1 - betas).cumprod(dim=0)
    if schedule == "sigmoid":
        start, end, tau = kw.get("start", -3), kw.get("end", 3), kw.get("tau", 1)
        v0, v1 = torch.tensor(start / tau).sigmoid(), torch.tensor(end / tau).sigmoid()
        f = (v1 - ((_grid(timesteps) * (end - start) + start) / tau).sigmoid()) / (v1 - v0)
        return (f / f[0])[1:]
    if schedule == "sd":
        betas = torch.linspace(kw.get("start", 0.00085) ** 0.5, kw.get("end", 0.0120) ** 0.5, timesteps,
                               dtype=torch.float64) ** 2
        return (1 - betas).cumprod(dim=0)
    raise ValueError(f"unknown beta schedule {schedule}")


def make_beta_schedule(schedule: str, timesteps: int, shift: float = 1.0, clip_min: float = 1e-9,
                       zero_terminal_snr: bool = True, **kw) -> torch.Tensor:
    ac = alphas_cumprod_for(schedule, timesteps, **kw)
    cosine_family = schedule in ("cosine", "cosine_simple_diffusion")
    if zero_terminal_snr and not cosine_family:
        root = torch.sqrt(ac)
        first, last = root[0].clone(), root[-1].clone()
        root = (root - last) * (first / (first - last))
        ac = root ** 2
        assert ac[-1] == 0, "terminal SNR not zero"
    if shift != 1.0 and schedule != "cosine_simple_diffusion":
        k = shift ** 2
        ac = k * ac / (k * ac + 1 - ac)
    alphas = torch.cat([ac[:1], ac[1:] / ac[:-1]])
    return torch.clip(1 - alphas, clip_min, 1.0)
