"""BaseVideoAlgo — the slice of the reference's algorithms/common/base_pytorch_video_algo.py that the
denoising sampling path uses: shape / latent / token bookkeeping (:37-88, :986-1033), model construction
(:142-174), normalisation (:489-502), condition processing (:635-700), scheduling matrices (:877-947) and
the validation entry point (:234-264), plus the VideoVAE decode of sampled latents (:507-629).
Training, VAE encoding, image VAEs, metrics and logging are out of scope: the hooks exist where callers
need them and raise or no-op explicitly.

It is a plain ``nn.Module`` (Lightning is not a dependency of the sampling path); ``state_dict()`` keys are
identical to the reference's (``data_mean``, ``data_std``, ``diffusion_model.model.*``).
"""
from typing import Callable, Dict, Optional

import numpy as np
import torch
from torch import Tensor, nn

from dfot_b200.checkpoint_io import load_checkpoint_file
from dfot_b200.config import to_config


class BaseVideoAlgo(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        cfg = to_config(cfg)
        self.cfg = cfg
        self.debug = cfg.get("debug", False)
        # 1. shape
        self.x_shape = list(cfg.x_shape)
        self.frame_skip = cfg.frame_skip
        self.chunk_size = cfg.chunk_size
        self.external_cond_type = cfg.external_cond_type
        self.external_cond_num_classes = cfg.external_cond_num_classes
        self.external_cond_dim = cfg.external_cond_dim * (cfg.frame_skip if cfg.external_cond_stack else 1)
        # 2. latent
        lat = cfg.latent
        self.is_latent_diffusion = lat.enabled
        self.is_latent_online = lat.type == "online"
        self.temporal_downsampling_factor = lat.downsampling_factor[0]
        self.is_latent_video_vae = self.temporal_downsampling_factor > 1
        if self.is_latent_diffusion:
            self.x_shape = list(lat.shape) if lat.get("shape") is not None else \
                [lat.num_channels] + [d // lat.downsampling_factor[1] for d in self.x_shape[1:]]
        # 3. diffusion
        d = cfg.diffusion
        self.use_causal_mask = d.use_causal_mask
        self.timesteps = d.timesteps
        self.sampling_timesteps = d.sampling_timesteps
        self.clip_noise = d.clip_noise
        self.is_full_sequence = (cfg.noise_level == "random_uniform" and not cfg.fixed_context.enabled
                                 and not cfg.variable_context.enabled)
        # 4. tasks
        self.logging = cfg.get("logging")
        self.tasks = [t for t in ("prediction", "interpolation") if cfg.tasks[t].enabled]
        self.generator = None
        self._build_model()

    # ------------------------------------------------------------------ construction
    def _build_model(self, diffusion_cls: Optional[Callable] = None) -> None:
        # `compile` (false / true / true_without_ddp_optimizer — the RE10K tree sets the last one) asks the reference to wrap
        # its PyTorch backbone in torch.compile; the backbones here are hand-written kernels replayed from CUDA graphs, so
        # the knob has nothing to act on and is accepted as it is
        self.diffusion_model = diffusion_cls(
            cfg=self.cfg.diffusion, backbone_cfg=self.cfg.backbone, x_shape=self.x_shape, max_tokens=self.max_tokens,
            external_cond_type=self.external_cond_type, external_cond_num_classes=self.external_cond_num_classes,
            external_cond_dim=self.external_cond_dim)
        self.register_buffer("data_mean", torch.tensor(self.cfg.data_mean).float())
        self.register_buffer("data_std", torch.tensor(self.cfg.data_std).float())
        self.vae = None

    @property
    def device(self) -> torch.device:
        return self.data_mean.device

    # ------------------------------------------------------------------ data plumbing (pre/post-processing, not per step)
    def on_after_batch_transfer(self, batch: Dict, dataloader_idx: int = 0) -> Dict:
        if self.is_latent_diffusion:
            if self.is_latent_online:
                raise NotImplementedError("online VAE encoding is outside the sampling-path scope; pass `latents`")
            xs = batch["latents"]
        else:
            xs = batch["videos"]
        gt_videos = batch.get("videos") if self.is_latent_diffusion else None
        xs = self._normalize_x(xs)
        if "masks" in batch:
            assert not self.is_latent_video_vae, "Masks should not be provided from the dataset when using VideoVAE."
            masks = batch["masks"]
        else:
            masks = torch.ones(*xs.shape[:2], dtype=torch.bool, device=xs.device)
        return {"xs": xs, "conditions": batch.get("conds", None), "masks": masks, "gt_videos": gt_videos}

    def _stat(self, t: Tensor, xs: Tensor) -> Tensor:
        return t.reshape([1] * (xs.ndim - t.ndim) + list(t.shape))

    def _normalize_x(self, xs: Tensor) -> Tensor:
        return (xs - self._stat(self.data_mean, xs)) / self._stat(self.data_std, xs)

    def _unnormalize_x(self, xs: Tensor) -> Tensor:
        return xs * self._stat(self.data_std, xs) + self._stat(self.data_mean, xs)

    # ------------------------------------------------------------------ latent decode (:507-629)
    def _load_vae(self) -> None:
        """(:507-551) — VAEs on the B200 kernels: the DC-AE image autoencoder when `vae.name` names it (DMLab / Minecraft),
        else the reference's causal VideoVAE (temporally compressed latents) or its ImageVAE; the diffusers KL autoencoder
        and TiTok tokenizer (external model code no shipped configuration of the sampling path selects) are not built."""
        name = self.cfg.vae.get("name")
        if name is not None and "dc_ae" in name:
            from ..vae import MyAutoencoderDC
            self.vae = MyAutoencoderDC.from_pretrained(cfg=self.cfg.vae, **dict(self.cfg.vae.get("pretrained_kwargs") or {}))
            self.vae = self.vae.to(self.device)
            for p in self.vae.parameters():
                p.requires_grad_(False)
            return
        if name is not None:
            raise NotImplementedError(f"vae.name={name!r} (diffusers AutoencoderKL / TiTok: external model code) is not "
                                      "decoded by dfot_b200; DC-AE and the reference's VideoVAE / ImageVAE are")
        from ..vae import ImageVAE, VideoVAE
        vae_cls = VideoVAE if self.is_latent_video_vae else ImageVAE
        self.vae = vae_cls.from_pretrained(path=self.cfg.vae.pretrained_path,
                                           **dict(self.cfg.vae.get("pretrained_kwargs") or {})).to(self.device)
        for p in self.vae.parameters():
            p.requires_grad_(False)

    @torch.no_grad()
    def _run_vae(self, x: Tensor, shape: str, vae_fn: Callable[[Tensor], Tensor]) -> Tensor:
        """(:555-585) — `shape` is a permutation of "b t c h w"; the batch is cut into cfg.vae.batch_size chunks; an
        image VAE sees the frames of a chunk as a batch of images."""
        axes = shape.split()
        x = x.permute(*[axes.index(a) for a in "bcthw"])
        n, step = x.shape[0], self.cfg.vae.batch_size
        outs = []
        for c in torch.chunk(x, (n + step - 1) // step, 0):
            if self.is_latent_video_vae:
                outs.append(vae_fn(c.contiguous()))
            else:
                b, ch, t, h, w = c.shape
                y = vae_fn(c.permute(0, 2, 1, 3, 4).reshape(b * t, ch, h, w))
                outs.append(y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4))
        y = torch.cat(outs, 0)
        return y.permute(*["bcthw".index(a) for a in axes])

    def _encode(self, x: Tensor, shape: str = "b t c h w") -> Tensor:
        raise NotImplementedError("VAE encoding is outside the scope of dfot_b200 (pass offline `latents`)")

    def _decode(self, latents: Tensor, shape: str = "b t c h w") -> Tensor:
        """(:599-629) — latent tokens -> frames in [0, 1]."""
        if self.vae is None:
            self._load_vae()
        if not self.is_latent_video_vae:
            return self._run_vae(latents, shape, lambda y: self.vae.decode(y) * 0.5 + 0.5)
        n_frames = self._n_tokens_to_n_frames(latents.shape[shape.split().index("t")])
        return self._run_vae(latents, shape, lambda y: self.vae.decode(y, n_frames) * 0.5 + 0.5)

    @torch.no_grad()
    def new_validation_step(self, batch, batch_idx, accelerator=None, namespace="validation", validate_sample=True):
        """(:234-264) — the denoising-loss evaluation is a training diagnostic and returns None here."""
        all_videos = None
        if validate_sample:
            all_videos = self._sample_all_videos(batch, batch_idx, namespace, n_context_tokens=self.n_context_tokens)
        return None, all_videos

    @torch.no_grad()
    def validation_step(self, batch, batch_idx, namespace="validation"):
        return self._sample_all_videos(batch, batch_idx, namespace)

    # ------------------------------------------------------------------ conditions / padding (:635-700)
    @torch.no_grad()
    def _process_conditions(self, conditions: Optional[Tensor], noise_levels: Optional[Tensor] = None):
        if conditions is None:
            return conditions
        mode = self.cfg.external_cond_processing
        if mode is None:
            return conditions
        if mode == "mask_first":
            out = conditions.clone()
            out[:, :1, : self.external_cond_dim] = 0
            return out
        raise NotImplementedError(f"External condition processing {mode} is not implemented.")

    def _pad_to_max_tokens(self, y: Optional[Tensor]) -> Optional[Tensor]:
        if y is None or y.shape[1] >= self.max_tokens:
            return y
        tail = y[:, -1:].expand(-1, self.max_tokens - y.shape[1], *y.shape[2:])
        return torch.cat([y, tail], dim=1)

    def _extend_x_dim(self, x: Tensor) -> Tensor:
        return x.reshape(*x.shape, *([1] * len(self.x_shape)))

    # ------------------------------------------------------------------ scheduling matrices (:877-947), host side
    def _generate_scheduling_matrix(self, horizon: int, padding: int = 0) -> Tensor:
        kind, S = self.cfg.scheduling_matrix, self.sampling_timesteps
        if kind in ("full_sequence", "gibbs"):
            idx = np.repeat(np.arange(S, -1, -1)[:, None], horizon, axis=1)
        elif kind == "autoregressive":
            idx = self._generate_pyramid_scheduling_matrix(horizon, S)
        elif kind == "interleaved":
            idx = self._generate_interleaved_scheduling_matrix(horizon, 3, S)
        else:
            raise ValueError(f"unknown scheduling matrix {kind}")
        levels = self.diffusion_model.ddim_idx_to_noise_level(torch.from_numpy(idx).long())
        if kind == "gibbs":   # one frame advances at a time; later frames hold the previous sweep's value
            sweeps = levels.shape[0]
            levels = levels.repeat_interleave(horizon, dim=0)
            for i in range(1, sweeps):
                prev_last = levels[(i - 1) * horizon + horizon - 1].clone()
                for j in range(horizon):
                    levels[i * horizon + j, j + 1:] = prev_last[j + 1:]
        return torch.nn.functional.pad(levels, (0, padding, 0, 0), value=self.timesteps - 1)

    def _generate_refine_scheduling_matrix(self, horizon: int, goback_length: int, n_goback: int,
                                           padding: int = 0) -> Tensor:
        """(:949-976, fork-only refinement sampling) the DDIM index walk S .. 0 which, after every index t in
        range(1, S - goback_length, goback_length), climbs back to t + goback_length and descends to t again, n_goback
        times; indices -> noise levels; pad columns are pure noise."""
        assert self.cfg.scheduling_matrix == "full_sequence", "Refining only support full_sequence scheduling matrix"
        S = self.sampling_timesteps
        goback = set(range(1, S - goback_length, goback_length))
        walk = []
        for t in range(S, -1, -1):
            walk.append(t)
            if t in goback:
                for _ in range(n_goback):
                    walk.extend(range(t + 1, t + goback_length + 1))
                    walk.extend(range(t + goback_length - 1, t - 1, -1))
        levels = self.diffusion_model.ddim_idx_to_noise_level(torch.tensor(walk).long())[:, None].repeat(1, horizon)
        return torch.nn.functional.pad(levels, (0, padding, 0, 0), value=self.timesteps - 1)

    def _generate_interleaved_scheduling_matrix(self, horizon: int, interleaved_size: int = 2,
                                                sampling_timesteps: int = 50) -> np.ndarray:
        S, k = sampling_timesteps, interleaved_size
        rows = S + k
        out = np.zeros((rows, horizon), dtype=np.int64)
        for t in range(horizon):
            lead = t % k + 1
            col = [S] * lead
            j = 0
            while len(col) < rows:
                idx = max(S - lead - k * j, 0)
                col += [idx] * (k if idx > 0 else rows - len(col))
                j += 1
            out[:, t] = col[:rows]
        return out

    def _generate_pyramid_scheduling_matrix(self, horizon: int, sampling_timesteps: int,
                                            uncertainty_scale: float = 1.0) -> np.ndarray:
        height = sampling_timesteps + int((horizon - 1) * uncertainty_scale) + 1
        lag = np.array([int(t * uncertainty_scale) for t in range(horizon)], dtype=np.int64)
        m = np.arange(height, dtype=np.int64)[:, None]
        return np.clip(sampling_timesteps + lag[None, :] - m, 0, sampling_timesteps)

    # ------------------------------------------------------------------ checkpoint ingestion (:1096-1201)
    # Same rules as the reference's Lightning hooks, so a reference `.ckpt` (full training state or the lightweight
    # EMA-only release files such as DFoT_RE10K.ckpt) or a `.safetensors` dump drops in: torch.compile prefixes are
    # stripped, EMA weights replace the raw ones for inference, only `diffusion_model.model.*` is taken from the file,
    # and missing keys raise unless `checkpoint.strict` is off.
    def _should_include_in_checkpoint(self, key: str) -> bool:
        return key.startswith("diffusion_model.model") or key.startswith("diffusion_model._orig_mod.model")

    def _load_ema_weights_to_state_dict(self, checkpoint: Dict) -> None:
        if checkpoint.get("pretrained_ema", False) and len(checkpoint.get("optimizer_states", [])) == 0:
            return   # EMA-only release checkpoint: the state_dict already holds the EMA weights
        ema_weights = checkpoint["optimizer_states"][0]["ema"]
        keys = ["diffusion_model." + k for k, _ in self.diffusion_model.named_parameters()]
        assert len(keys) == len(ema_weights), "Number of original weights and EMA weights do not match."
        for key, weight in zip(keys, ema_weights):
            checkpoint["state_dict"][key] = weight

    def on_save_checkpoint(self, checkpoint: Dict) -> None:
        state_dict = checkpoint["state_dict"]
        for key in list(state_dict.keys()):
            if not self._should_include_in_checkpoint(key):
                del state_dict[key]

    def on_load_checkpoint(self, checkpoint: Dict) -> None:
        sd = checkpoint["state_dict"]
        # a checkpoint written by a torch.compile'd reference model carries `_orig_mod.` in its keys (:1099-1113)
        checkpoint["state_dict"] = sd = {k.replace("diffusion_model._orig_mod.", "diffusion_model."): v for k, v in sd.items()}
        if "optimizer_states" in checkpoint and (len(checkpoint["optimizer_states"]) > 0 or
                                                 not checkpoint.get("pretrained_ema", False)):
            if len(checkpoint["optimizer_states"]) > 0 and "ema" in checkpoint["optimizer_states"][0]:
                self._load_ema_weights_to_state_dict(checkpoint)
        new_sd = {}
        own = self.state_dict()
        for key, value in own.items():
            new_sd[key] = sd[key] if self._should_include_in_checkpoint(key) and key in sd else value
        self.ckpt_ignored_keys = [k for k in sd if not self._should_include_in_checkpoint(k)]
        self.ckpt_missing_keys = [k for k in own if self._should_include_in_checkpoint(k) and k not in sd]
        if self.ckpt_missing_keys and self.cfg.checkpoint.strict:
            raise ValueError(f"The following keys are not found in the checkpoint: {self.ckpt_missing_keys}. Thus, the "
                             "checkpoint cannot be loaded. To ignore this error, turn off strict checkpoint loading by "
                             "setting `algorithm.checkpoint.strict=False`.")
        checkpoint["state_dict"] = new_sd

    def load_checkpoint(self, path: str) -> None:
        """Load model weights from a reference checkpoint file (`.ckpt` Lightning / Accelerate dict or `.safetensors`,
        experiments/simple_video_generation.py:602-629) through the hooks above."""
        if path.endswith(".safetensors"):
            from safetensors.torch import load_file
            checkpoint = {"state_dict": load_file(path), "pretrained_ema": True, "optimizer_states": []}
        else:
            checkpoint = load_checkpoint_file(path)
            if "state_dict" not in checkpoint:
                checkpoint = {"state_dict": checkpoint, "pretrained_ema": True, "optimizer_states": []}
        self.on_load_checkpoint(checkpoint)
        self.load_state_dict(checkpoint["state_dict"], strict=True)

    # ------------------------------------------------------------------ frames vs tokens (:986-1033)
    def _n_frames_to_n_tokens(self, n_frames: int) -> int:
        return (n_frames - 1) // self.temporal_downsampling_factor + 1

    def _n_tokens_to_n_frames(self, n_tokens: int) -> int:
        return (n_tokens - 1) * self.temporal_downsampling_factor + 1

    @property
    def max_frames(self) -> int:
        return self.cfg.max_frames

    @property
    def max_tokens(self) -> int:
        return self._n_frames_to_n_tokens(self.max_frames)

    @property
    def n_frames(self) -> int:
        return self.cfg.n_frames

    @property
    def n_context_frames(self) -> int:
        return self.cfg.context_frames

    @property
    def n_tokens(self) -> int:
        return self._n_frames_to_n_tokens(self.n_frames)

    @property
    def n_context_tokens(self) -> int:
        return self._n_frames_to_n_tokens(self.n_context_frames)
