"""DiscreteDiffusion — sampling-time drop-in for the reference's
algorithms/dfot/diffusion/discrete_diffusion.py:32-550 (same constructor, buffer names, ``q_sample``,
``ddim_idx_to_noise_level``, ``sample_step``, ``model_predictions``).

The per-frame DDIM algebra runs in the fused K4 kernel; this class owns the float tables (built on the
host in float64 exactly like the reference, registered as non-persistent fp32 buffers under the same
names) and the backbone.  Training (``forward``, loss weighting) is out of scope.
"""
from collections import namedtuple
from typing import Callable, Optional

import numpy as np
import torch
from torch import nn

from dfot_b200 import ops
from dfot_b200.config import to_config
from .. import sampling_plan as sp
from ..backbones import DiT3D, UViT3DPose
from .noise_schedule import make_beta_schedule

ModelPrediction = namedtuple("ModelPrediction", ["pred_noise", "pred_x_start", "model_out"])


class DiscreteDiffusion(nn.Module):
    is_continuous = False
    precond_scale = 1.0

    def __init__(self, cfg, backbone_cfg, x_shape, max_tokens: int, external_cond_type, external_cond_num_classes,
                 external_cond_dim: int):
        super().__init__()
        cfg, backbone_cfg = to_config(cfg), to_config(backbone_cfg)
        self.cfg = cfg
        self.x_shape = x_shape
        self.max_tokens = max_tokens
        self.external_cond_type = external_cond_type
        self.external_cond_num_classes = external_cond_num_classes
        self.external_cond_dim = external_cond_dim
        self.timesteps = cfg.timesteps
        self.sampling_timesteps = cfg.sampling_timesteps
        self.beta_schedule = cfg.beta_schedule
        self.schedule_fn_kwargs = cfg.schedule_fn_kwargs
        self.objective = cfg.objective
        self.loss_weighting = cfg.loss_weighting
        self.ddim_sampling_eta = cfg.ddim_sampling_eta
        self.clip_noise = cfg.clip_noise
        self.backbone_cfg = backbone_cfg
        self.use_causal_mask = cfg.use_causal_mask
        self.noise_source: Optional[Callable] = None   # test hook: callable(shape, device) -> standard normal tensor
        self.generator: Optional[torch.Generator] = None   # per-shard noise stream (DFoTVideo.sample_sharded); None = global
        self._skip_increment = {}
        self._build_model()
        self._build_buffer()

    # discrete_diffusion.py:64-92
    def _build_model(self):
        name = self.backbone_cfg.name
        if name == "dit3d":
            model_cls = DiT3D
        elif name == "u_vit3d_pose":
            model_cls = UViT3DPose
        elif name in ("u_net3d", "u_vit3d", "dit3d_pose", "far_dit", "dit1d", "difference_dit3d"):
            raise NotImplementedError(f"backbone `{name}` is not implemented by dfot_b200 yet (see DESIGN.md scope)")
        else:
            raise ValueError(f"unknown model type {name}")
        self.model = model_cls(cfg=self.backbone_cfg, x_shape=self.x_shape, max_tokens=self.max_tokens,
                               external_cond_type=self.external_cond_type,
                               external_cond_num_classes=self.external_cond_num_classes,
                               external_cond_dim=self.external_cond_dim, use_causal_mask=self.use_causal_mask)

    # discrete_diffusion.py:94-168
    def _build_buffer(self):
        betas = make_beta_schedule(schedule=self.beta_schedule, timesteps=self.timesteps,
                                   zero_terminal_snr=self.objective != "pred_noise",
                                   **dict(self.schedule_fn_kwargs or {}))
        alphas = 1.0 - betas
        ac = torch.cumprod(alphas, dim=0)
        ac_prev = torch.cat([torch.ones(1, dtype=ac.dtype), ac[:-1]])
        assert self.sampling_timesteps <= self.timesteps
        self.is_ddim_sampling = self.sampling_timesteps < self.timesteps
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        snr = ac / (1 - ac)
        tables = {
            "betas": betas, "alphas_cumprod": ac, "alphas_cumprod_prev": ac_prev,
            "sqrt_alphas_cumprod": torch.sqrt(ac), "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
            "log_one_minus_alphas_cumprod": torch.log(1.0 - ac),
            "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / ac), "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / ac - 1),
            "posterior_variance": post_var, "posterior_log_variance_clipped": torch.log(post_var.clamp(min=1e-20)),
            "posterior_mean_coef1": betas * torch.sqrt(ac_prev) / (1.0 - ac),
            "posterior_mean_coef2": (1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac), "snr": snr,
        }
        strategy = self.loss_weighting.strategy
        if strategy in {"min_snr", "fused_min_snr"}:
            tables["clipped_snr"] = snr.clamp(max=self.loss_weighting.snr_clip)
        elif strategy == "sigmoid":
            tables["logsnr"] = torch.log(snr)
        for name, val in tables.items():
            self.register_buffer(name, val.to(torch.float32), persistent=False)
        f = lambda n: tables[n].to(torch.float32).numpy().copy()
        self.host_tables = sp.HostTables(
            alphas_cumprod=f("alphas_cumprod"), sqrt_alphas_cumprod=f("sqrt_alphas_cumprod"),
            sqrt_one_minus_alphas_cumprod=f("sqrt_one_minus_alphas_cumprod"),
            sqrt_recip_alphas_cumprod=f("sqrt_recip_alphas_cumprod"),
            sqrt_recipm1_alphas_cumprod=f("sqrt_recipm1_alphas_cumprod"),
            logsnr=f("logsnr") if "logsnr" in tables else None, objective=self.objective,
            eta=float(self.ddim_sampling_eta), clip_noise=float(self.clip_noise), timesteps=self.timesteps,
            is_ddim=self.is_ddim_sampling, posterior_mean_coef1=f("posterior_mean_coef1"),
            posterior_mean_coef2=f("posterior_mean_coef2"),
            posterior_log_variance_clipped=f("posterior_log_variance_clipped"))

    # ------------------------------------------------------------------ noise plumbing
    def randn(self, shape, device) -> torch.Tensor:
        """Standard normal noise in the reference's draw order (SURVEY.md §8a RNG contract): torch's generator
        for ``device`` unless a test injected ``noise_source``."""
        if self.noise_source is not None:
            return self.noise_source(tuple(shape), device)
        return torch.randn(tuple(shape), device=device, generator=self.generator)

    def clipped_noise(self, shape, device) -> torch.Tensor:
        return torch.clamp(self.randn(shape, device), -self.clip_noise, self.clip_noise)

    def skip_randn(self, shape, device) -> None:
        """Advance the noise stream exactly as `randn(shape, device)` would, without producing the values.  The reference
        draws the per-step update noise even when it is multiplied by sigma = 0 (DDIM, eta = 0: discrete_diffusion.py:525),
        and the lockstep rounds replay whole windows only to find stream positions; both want the position, not the numbers.
        CUDA generators are counter based: the Philox offset a draw of a given size consumes is measured once and re-applied."""
        if self.noise_source is not None or torch.device(device).type != "cuda":
            self.randn(shape, device)
            return
        g = self._torch_generator(device)
        key = (int(np.prod(shape)), str(device))
        inc = self._skip_increment.get(key)
        if inc is None:
            before = g.get_offset()
            torch.randn(tuple(shape), device=device, generator=self.generator)
            self._skip_increment[key] = g.get_offset() - before
        else:
            g.set_offset(g.get_offset() + inc)

    def _torch_generator(self, device) -> torch.Generator:
        if self.generator is not None:
            return self.generator
        device = torch.device(device)
        if device.type == "cuda":
            return torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
        return torch.default_generator

    def noise_get_state(self, device):
        """Position of the noise stream (host-side: seed + offset of torch's generator, or the injected source's own state).
        Lets independent windows that the reference samples one after the other advance TOGETHER (lockstep rounds,
        DFoTVideo._run_lockstep) while each still draws exactly the values the sequential order gives it."""
        if self.noise_source is not None:
            get = getattr(self.noise_source, "get_state", None)
            if get is None:
                raise RuntimeError("the injected noise_source cannot fork its stream (no get_state / set_state)")
            return get()
        return self._torch_generator(device).get_state()

    def noise_set_state(self, device, state) -> None:
        if self.noise_source is not None:
            self.noise_source.set_state(state)
        else:
            self._torch_generator(device).set_state(state)

    def noise_can_fork(self) -> bool:
        return self.noise_source is None or (hasattr(self.noise_source, "get_state") and hasattr(self.noise_source, "set_state"))

    # ------------------------------------------------------------------ reference API
    def ddim_idx_to_noise_level(self, indices: torch.Tensor) -> torch.Tensor:
        # fp32 linspace then truncation (discrete_diffusion.py:379-384) — bit-exact host arithmetic
        real_steps = torch.linspace(-1, self.timesteps - 1, self.sampling_timesteps + 1).long().to(indices.device)
        return real_steps[indices.flatten()].view(indices.shape)

    def q_sample(self, x_start: torch.Tensor, k: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """sqrt(ᾱ_k)·x0 + sqrt(1-ᾱ_k)·noise per frame (:242-250) — one K4 launch in prepare-only mode."""
        R, T = k.shape
        if noise is None:
            noise = self.clipped_noise(x_start.shape, x_start.device)
        if tuple(noise.shape) != tuple(x_start.shape) or tuple(x_start.shape[:2]) != (R, T):
            raise ValueError(f"q_sample: x_start {tuple(x_start.shape)}, k {tuple(k.shape)} and noise "
                             f"{tuple(noise.shape)} do not describe the same (rows, frames, ...) batch")
        kh = k.detach().cpu().numpy().astype(np.int64)
        prep = np.zeros((R, T), dtype=sp.PREPARE_DTYPE)
        prep["mode"] = sp.MODE_QSAMPLE
        prep["noise_row"] = np.arange(R, dtype=np.int32)[:, None]
        prep["qa"] = self.host_tables.sqrt_alphas_cumprod[kh]
        prep["qb"] = self.host_tables.sqrt_one_minus_alphas_cumprod[kh]
        out = torch.empty(x_start.shape, dtype=torch.float32, device=x_start.device)
        ops.sampler_step_hg(x_start.contiguous().float().clone(), None, out, None,
                            sp.to_device_bytes(prep, x_start.device), None, noise.contiguous().float(), None, R, 1, T,
                            max_noise_row=R - 1)
        return out

    def renoise_table(self, cur: np.ndarray, nxt: np.ndarray) -> np.ndarray:
        """Per-frame records of `q_sample_from_x_k` (host): qa = sqrt(s), qb = sqrt(1 - s), s = ᾱ[next] / ᾱ[cur] in fp32
        (index -1 = last entry, as in the reference), s = 1 where next == 999."""
        ac = self.host_tables.alphas_cumprod.astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            scale = (ac[nxt] / ac[cur]).astype(np.float32)
            scale = np.where(nxt == 999, np.float32(1.0), scale)
            prep = np.zeros(cur.shape, dtype=sp.PREPARE_DTYPE)
            prep["mode"] = sp.MODE_QSAMPLE
            prep["noise_row"] = np.arange(cur.shape[0], dtype=np.int32)[:, None]
            prep["qa"] = np.sqrt(scale)
            prep["qb"] = np.sqrt(np.float32(1.0) - scale)
        return prep

    def q_sample_from_x_k(self, x_k: torch.Tensor, cur_noise_levels: torch.Tensor, next_noise_levels: torch.Tensor,
                          noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(:252-260, refinement sampling) forward diffusion from level `cur` up to level `next`:
        sqrt(s)·x_k + sqrt(1-s)·noise — one K4 launch in prepare-only mode."""
        R, T = cur_noise_levels.shape
        if noise is None:
            noise = self.clipped_noise(x_k.shape, x_k.device)
        if tuple(noise.shape) != tuple(x_k.shape) or tuple(x_k.shape[:2]) != (R, T):
            raise ValueError(f"q_sample_from_x_k: x_k {tuple(x_k.shape)}, levels {(R, T)} and noise {tuple(noise.shape)} "
                             "do not describe the same (rows, frames, ...) batch")
        prep = self.renoise_table(cur_noise_levels.detach().cpu().numpy().astype(np.int64),
                                  next_noise_levels.detach().cpu().numpy().astype(np.int64))
        out = torch.empty(x_k.shape, dtype=torch.float32, device=x_k.device)
        ops.sampler_step_hg(x_k.contiguous().float().clone(), None, out, None, sp.to_device_bytes(prep, x_k.device), None,
                            noise.contiguous().float(), None, R, 1, T, max_noise_row=R - 1)
        return out

    def model_input_levels(self, k: torch.Tensor) -> torch.Tensor:
        return k

    def model_predictions(self, x, k, external_cond=None, external_cond_mask=None) -> ModelPrediction:
        """(:173-223) — API parity; the sampler itself never materialises x0 / eps."""
        out = self.model(x, self.model_input_levels(k), external_cond, external_cond_mask).clone()
        R, T = k.shape
        kh = k.detach().cpu().numpy().astype(np.int64)
        tb = self.host_tables
        sa, sb = tb.sqrt_alphas_cumprod[kh].astype(np.float64), tb.sqrt_one_minus_alphas_cumprod[kh].astype(np.float64)
        if self.objective == "pred_v":
            coef = {"x0": (sa, -sb, 0.0), "eps": (sb, sa, 0.0)}
        elif self.objective == "pred_x0":
            coef = {"x0": (0 * sa, 1 + 0 * sa, 0.0), "eps": (1 / sb, -sa / sb, 0.0)}
        elif self.objective == "pred_noise":
            sr, srm1 = tb.sqrt_recip_alphas_cumprod[kh].astype(np.float64), tb.sqrt_recipm1_alphas_cumprod[kh].astype(np.float64)
            coef = {"x0": (sr, -srm1, self.clip_noise), "eps": (0 * sa, 1 + 0 * sa, self.clip_noise)}
        else:
            raise ValueError(f"unknown objective {self.objective}")
        res = {}
        for name, (a, b, clip) in coef.items():
            upd = np.zeros((R, T), dtype=sp.UPDATE_DTYPE)
            upd["a"], upd["b"], upd["w"], upd["clip"], upd["generate"] = a, b, 1.0, clip, 1
            y = x.contiguous().float().clone()
            ops.sampler_step_hg(y, out, None, sp.to_device_bytes(upd, x.device), None, None, None, None, R, 1, T)
            res[name] = y
        return ModelPrediction(res["eps"], res["x0"], out)

    def sample_step(self, x: torch.Tensor, curr_noise_level: torch.Tensor, next_noise_level: torch.Tensor,
                    external_cond: Optional[torch.Tensor], external_cond_mask: Optional[torch.Tensor] = None,
                    guidance_fn: Optional[Callable] = None) -> torch.Tensor:
        """(:386-538) backbone forward + fused per-frame update: DDIM, or the DDPM ancestral step (:423-452) when
        sampling_timesteps == timesteps — the same kernel with another host coefficient table.  ``guidance_fn``
        (reconstruction guidance) needs autograd through the backbone and is out of scope."""
        if guidance_fn is not None:
            raise NotImplementedError("guidance_fn / reconstruction guidance is not supported by dfot_b200")
        R, T = curr_noise_level.shape
        frm = curr_noise_level.detach().cpu().numpy().astype(np.int64)
        to = next_noise_level.detach().cpu().numpy().astype(np.int64)
        tb = self.host_tables
        upd = sp.step_update_table(tb, frm, to, np.ones((R, T), np.float32), np.ones((R, T), np.int32))
        levels = torch.from_numpy(sp.model_levels(tb, frm, self.is_continuous, self.precond_scale)).to(x.device)
        out = self.model(x, levels, external_cond, external_cond_mask)
        noise = self.clipped_noise(x.shape, x.device)          # drawn even when eta == 0 (:525-526)
        y = x.contiguous().float().clone()
        ops.sampler_step_hg(y, out, None, sp.to_device_bytes(upd, x.device), None,
                            noise if tb.uses_step_noise else None, None, None, R, 1, T)
        return y
