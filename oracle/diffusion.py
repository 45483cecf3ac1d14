"""ORACLE (test infrastructure): per-frame diffusion algebra.

Restates algorithms/dfot/diffusion/discrete_diffusion.py
  q_sample :242-250, model_predictions :173-191, predict_* :193-223,
  ddim_sample_step :454-538, ddpm_sample_step :423-452 (+ q_posterior :231-240)
and continuous_diffusion.py:118-138 (model is fed precond_scale*logsnr[k]).
"""
from typing import Callable, Dict, Optional

import torch

from .schedule import diffusion_buffers


def _per_frame(a: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return a.reshape(*a.shape, *([1] * (like.ndim - a.ndim)))


class Diffusion:
    """Functional stand-in for DiscreteDiffusion / ContinuousDiffusion at sampling time.

    ``model``: callable (x, noise_levels, external_cond, external_cond_mask) -> tensor like x.
    ``randn_like``: noise source (lets tests share one noise stream between oracle and product).
    """

    def __init__(self, diff_cfg: dict, model: Callable, randn_like: Callable = torch.randn_like):
        self.cfg = diff_cfg
        self.model = model
        self.randn_like = randn_like
        self.buf: Dict[str, torch.Tensor] = diffusion_buffers(diff_cfg)
        self.timesteps = diff_cfg["timesteps"]
        self.sampling_timesteps = diff_cfg["sampling_timesteps"]
        self.objective = diff_cfg["objective"]
        self.eta = diff_cfg["ddim_sampling_eta"]
        self.clip_noise = diff_cfg["clip_noise"]
        self.is_continuous = bool(diff_cfg.get("is_continuous", False))
        self.precond_scale = diff_cfg.get("precond_scale", 1.0)
        self.is_ddim_sampling = self.sampling_timesteps < self.timesteps       # else: DDPM ancestral step (:403-413)

    def clipped_noise(self, like: torch.Tensor) -> torch.Tensor:
        return torch.clamp(self.randn_like(like), -self.clip_noise, self.clip_noise)

    def q_sample(self, x0: torch.Tensor, k: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        # discrete_diffusion.py:242-250 (k = -1 indexes the last entry; callers mask it out)
        if noise is None:
            noise = self.clipped_noise(x0)
        a = _per_frame(self.buf["sqrt_alphas_cumprod"][k], x0)
        s = _per_frame(self.buf["sqrt_one_minus_alphas_cumprod"][k], x0)
        return a * x0 + s * noise

    def q_sample_from_x_k(self, x_k: torch.Tensor, cur: torch.Tensor, nxt: torch.Tensor,
                          noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        # discrete_diffusion.py:252-260 — forward diffusion from level `cur` up to level `nxt` (refinement sampling)
        if noise is None:
            noise = self.clipped_noise(x_k)
        ac = self.buf["alphas_cumprod"]
        scale = _per_frame(ac[nxt], x_k) / _per_frame(ac[cur], x_k)
        scale = torch.where(_per_frame(nxt, x_k) == 999, torch.ones_like(scale), scale)
        return torch.sqrt(scale) * x_k + torch.sqrt(1 - scale) * noise

    def model_input_level(self, k: torch.Tensor) -> torch.Tensor:
        # discrete: int64 k; continuous: fp32 precond_scale * logsnr[k]  (continuous_diffusion.py:118-121)
        return self.precond_scale * self.buf["logsnr"][k] if self.is_continuous else k

    def predictions(self, x, k, cond=None, cond_mask=None):
        # discrete_diffusion.py:173-223
        out = self.model(x, self.model_input_level(k), cond, cond_mask)
        sa = _per_frame(self.buf["sqrt_alphas_cumprod"][k], x)
        sb = _per_frame(self.buf["sqrt_one_minus_alphas_cumprod"][k], x)
        if self.objective == "pred_v":
            x0 = sa * x - sb * out
            eps = sa * out + sb * x
        elif self.objective == "pred_noise":
            eps = torch.clamp(out, -self.clip_noise, self.clip_noise)
            x0 = _per_frame(self.buf["sqrt_recip_alphas_cumprod"][k], x) * x \
                - _per_frame(self.buf["sqrt_recipm1_alphas_cumprod"][k], x) * eps
        elif self.objective == "pred_x0":
            x0 = out
            eps = (x - sa * x0) / sb
        else:
            raise ValueError(self.objective)
        return eps, x0, out

    def ddim_coefficients(self, curr: torch.Tensor, nxt: torch.Tensor):
        # discrete_diffusion.py:464-478 (fp32 tables, int64 ones/zeros promoted)
        k = torch.clamp(curr, min=0)
        ac = self.buf["alphas_cumprod"]
        alpha = ac[k]
        alpha_next = torch.where(nxt < 0, torch.ones_like(nxt), ac[nxt])
        sigma = torch.where(nxt < 0, torch.zeros_like(nxt),
                            self.eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt())
        c = (1 - alpha_next - sigma ** 2).sqrt()
        return k, alpha_next, sigma, c

    def ddpm_sample_step(self, x, curr, cond=None, cond_mask=None, return_model_out: bool = False):
        # discrete_diffusion.py:423-452: x0 from the model, posterior mean / variance at level k, fresh noise where k > 0
        k = torch.clamp(curr, min=0)
        _, x0, out = self.predictions(x, k, cond, cond_mask)
        mean = _per_frame(self.buf["posterior_mean_coef1"][k], x) * x0 + _per_frame(self.buf["posterior_mean_coef2"][k], x) * x
        log_var = _per_frame(self.buf["posterior_log_variance_clipped"][k], x)
        noise = torch.where(_per_frame(k > 0, x), self.randn_like(x), torch.zeros_like(x))
        noise = torch.clamp(noise, -self.clip_noise, self.clip_noise)
        x_new = mean + torch.exp(0.5 * log_var) * noise
        x_new = torch.where(_per_frame(curr == -1, x), x, x_new)
        return (x_new, out) if return_model_out else x_new

    def sample_step(self, x, curr, nxt, cond=None, cond_mask=None, return_model_out: bool = False):
        if not self.is_ddim_sampling:
            return self.ddpm_sample_step(x, curr, cond, cond_mask, return_model_out)
        # discrete_diffusion.py:454-538, guidance_fn=None branch
        k, alpha_next, sigma, c = self.ddim_coefficients(curr, nxt)
        eps, x0, out = self.predictions(x, k, cond, cond_mask)
        noise = self.clipped_noise(x)  # drawn even when eta == 0, :525-526
        x_new = x0 * _per_frame(alpha_next, x).sqrt() + eps * _per_frame(c, x) + _per_frame(sigma, x) * noise
        x_new = torch.where(_per_frame(curr == nxt, x), x, x_new)
        return (x_new, out) if return_model_out else x_new
