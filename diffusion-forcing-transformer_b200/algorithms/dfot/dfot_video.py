"""DFoTVideo — B200-native drop-in for the sampling half of the reference's
algorithms/dfot/dfot_video.py (same method names and signatures):
    _sample_all_videos :80-112, _predict_videos :114-179, _interpolate_videos :181-360,
    _predict_sequence :362-514, _sample_sequence :516-763.

Design (not a translation): a denoising window is *planned* on the host before its first step
(sampling_plan.py / HistoryGuidance.plan_step) — scheduling matrix, context-mask evolution, guidance
branch tables, DDIM coefficients and re-noising instructions for every (step, branch row, frame) — and
uploaded once.  A sampling step is then:  backbone forward (tcgen05 GEMMs + attention kernels)  →  ONE
fused K4 launch that applies the per-frame DDIM update, the history-guidance combine, the context revert
and emits the next step's branch inputs.  No device→host sync happens inside the loop.
Noise is drawn with torch's generator in the reference's order (SURVEY.md §8a RNG contract).
"""
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from dfot_b200 import ops
from dfot_b200.algorithms.common.base_pytorch_video_algo import BaseVideoAlgo
from . import sampling_plan as sp
from .diffusion import ContinuousDiffusion, DiscreteDiffusion
from .history_guidance import HistoryGuidance


def interpolation_plan(known: np.ndarray, max_tokens: int) -> List[List[np.ndarray]]:
    """Rounds of frame-index chunks that fill every unknown frame between known ones (dfot_video.py:219-261):
    gaps >= max_tokens-1 get max_tokens equally spaced frames; shorter neighbouring gaps are merged into one
    chunk while they fit in max_tokens."""
    known = np.asarray(known, dtype=bool).copy()
    rounds = []
    while not known.all():
        keys = np.nonzero(known)[0]
        chunks, pending = [], None
        for left, right in zip(keys[:-1], keys[1:]):
            gap = int(right - left)
            if pending is not None:
                if len(pending) + gap <= max_tokens:
                    pending = np.concatenate([pending, np.arange(left + 1, right + 1)])
                    continue
                chunks.append(pending)
                pending = None
            if gap == 1:
                continue
            if gap >= max_tokens - 1:
                chunks.append(torch.linspace(int(left), int(right), max_tokens).round().long().numpy())
            else:
                pending = np.arange(left, right + 1)
        if pending is not None:
            chunks.append(pending)
        for c in chunks:
            known[c] = True
        rounds.append(chunks)
    return rounds


class DFoTVideo(BaseVideoAlgo):
    def __init__(self, cfg):
        super().__init__(cfg)
        self.trace: Optional[list] = None       # tests: per-step tensors are appended when this is a list
        self.nfe_rows = 0                       # backbone forward-rows executed (NFE counter)
        self.nfe_rows_planned = 0               # ... including rows another shard executes (dry_run replays)
        self.model_in_dtype = torch.bfloat16    # dtype of the branch inputs emitted by K4
        self.mesh = None                        # dfot_b200.distributed.Mesh for multi-GPU sampling (None = 1 GPU)
        self.shard_chunks = False               # interpolation chunk batches spread over the dp axis (sample_sharded)

    def _build_model(self) -> None:
        super()._build_model(ContinuousDiffusion if self.cfg.diffusion.is_continuous else DiscreteDiffusion)

    def training_step(self, *a, **k):
        raise NotImplementedError("training is outside the scope of dfot_b200 (sampling path only)")

    # ------------------------------------------------------------------ entry points
    @torch.no_grad()
    def _sample_all_videos(self, batch, batch_idx=0, namespace="validation", n_context_tokens=None):
        xs, conditions = batch["xs"], batch.get("conditions")
        n_ctx = n_context_tokens if n_context_tokens is not None else self.n_context_tokens
        videos: Dict[str, Tensor] = {"gt": xs.clone()}
        for task in self.tasks:
            fn = self._predict_videos if task == "prediction" else self._interpolate_videos
            videos[task] = fn(xs, conditions=conditions, n_context_tokens=n_ctx)
        videos = {k: self._unnormalize_x(v).detach() for k, v in videos.items() if v is not None}
        if self.is_latent_diffusion and (self.vae is not None or (self.cfg.get("vae") or {}).get("pretrained_path")):
            # (:104-111) decode latents to frames; with no decoder configured (vae.pretrained_path null and no `vae`
            # attached) the latents themselves are returned
            gt = batch.get("gt_videos")
            videos = {k: (gt if k == "gt" and gt is not None else self._decode(v)) for k, v in videos.items()}
        return videos

    @torch.no_grad()
    def _predict_videos(self, xs: Tensor, n_context_tokens: int, conditions: Optional[Tensor] = None) -> Tensor:
        task = self.cfg.tasks.prediction
        guidance = HistoryGuidance.from_config(task.history_guidance, timesteps=self.timesteps)
        density = task.get("keyframe_density") or 1
        if density > 1:
            raise ValueError("tasks.prediction.keyframe_density must be <= 1")
        T = xs.shape[1]
        keys = torch.linspace(0, T - 1, round(density * T)).round().long()
        keys = torch.cat([torch.arange(n_context_tokens), keys]).unique()      # context frames are keyframes
        key_cond = None
        if conditions is not None:
            if self.external_cond_type == "label":
                key_cond = conditions
            elif self.external_cond_type == "action":
                key_cond = conditions[:, keys.to(conditions.device)]
            else:
                raise ValueError(f"Unknown external condition type: {self.external_cond_type}. "
                                 "Supported types are 'label' and 'action'.")
        out = xs.clone()
        pred, _ = self._predict_sequence(
            out[:, :n_context_tokens], length=len(keys), conditions=key_cond, history_guidance=guidance,
            reconstruction_guidance=self.cfg.diffusion.reconstruction_guidance,
            sliding_context_len=task.get("sliding_context_len") or self.max_tokens // 2)
        out[:, keys.to(out.device)] = pred.to(out.dtype)
        if len(keys) < T:
            known = torch.zeros(out.shape[:2], dtype=torch.bool)
            known[:, keys] = True
            out = self._interpolate_videos(context=out, context_mask=known, conditions=conditions)
        return out

    @torch.no_grad()
    def _interpolate_videos(self, context: Tensor, context_mask: Optional[Tensor] = None,
                            conditions: Optional[Tensor] = None, **kwargs) -> Tensor:
        B, T = context.shape[:2]
        if context_mask is None:
            known = np.zeros((B, T), dtype=bool)
            known[:, [0, -1]] = True
        else:
            known = context_mask.detach().cpu().numpy().astype(bool)
            assert known[:, [0, -1]].all(), "The first and last frames must be known to interpolate."
        task = self.cfg.tasks.interpolation
        guidance = HistoryGuidance.from_config(task.history_guidance, timesteps=self.timesteps)
        xs = context.clone()
        for chunks in interpolation_plan(known[0], self.max_tokens):
            idx = [torch.from_numpy(c).to(xs.device) for c in chunks]
            ctx = torch.cat([self._pad_to_max_tokens(xs[:, i]) for i in idx], 0)
            msk = np.concatenate([self._pad_rows(known[:, c]) for c in chunks], 0)
            cnd = None
            if conditions is not None:
                if self.external_cond_type == "label":
                    cnd = torch.cat([conditions for _ in idx], 0)
                elif self.external_cond_type == "action":
                    cnd = torch.cat([self._pad_to_max_tokens(conditions[:, i.to(conditions.device)]) for i in idx], 0)
                else:
                    raise ValueError(f"Unknown external condition type: {self.external_cond_type}. "
                                     "Supported types are 'label' and 'action'.")
            rows = ctx.shape[0]
            mb = task.get("max_batch_size") or rows
            # Chunk batches of one round are independent (SURVEY.md §8e axis 3).  With `shard_chunks` (set by
            # sample_sharded when the samples themselves cannot fill the dp axis) batch i is sampled by dp shard
            # i mod dp; the other shards replay its noise draws only, so the result equals the single-GPU rollout.
            mesh = self.mesh if self.shard_chunks else None
            outs, owners = [], []
            for bi, s in enumerate(range(0, rows, mb)):   # every chunk is processed (the reference's conditions=None
                e = min(rows, s + mb)                     # path drops the last partial batch — quirk Q10, not replicated)
                owner = bi % mesh.dp if mesh is not None else 0
                o, _ = self._window_sampler()(batch_size=e - s, context=ctx[s:e],
                                             context_mask=torch.from_numpy(msk[s:e].astype(np.int64)),
                                             conditions=None if cnd is None else cnd[s:e], history_guidance=guidance,
                                             dry_run=mesh is not None and owner != mesh.dp_index)
                outs.append(o)
                owners.append(owner)
            if mesh is not None:
                from dfot_b200 import distributed as D
                outs = [D.broadcast_from_shard(o.contiguous(), owner, mesh) for o, owner in zip(outs, owners)]
            outs = torch.cat(outs, 0)
            for c, i, pred in zip(chunks, idx, outs.chunk(len(chunks), 0)):
                xs[:, i] = pred[:, : len(c)]
                known[:, c] = True
        return xs

    def _pad_rows(self, m: np.ndarray) -> np.ndarray:
        if m.shape[1] >= self.max_tokens:
            return m
        return np.concatenate([m, np.repeat(m[:, -1:], self.max_tokens - m.shape[1], 1)], 1)

    @torch.no_grad()
    def _predict_sequence(self, context: Tensor, length: Optional[int] = None, conditions: Optional[Tensor] = None,
                          guidance_fn: Optional[Callable] = None, reconstruction_guidance: float = 0.0,
                          history_guidance: Optional[HistoryGuidance] = None,
                          sliding_context_len: Optional[int] = None, return_all: bool = False
                          ) -> Tuple[Tensor, Optional[Tensor]]:
        if length is None:
            length = self.max_tokens
        if sliding_context_len is None:
            if self.max_tokens < length:
                raise ValueError("when length > max_tokens, sliding_context_len must be specified.")
            sliding_context_len = self.max_tokens - 1
        if sliding_context_len == -1:
            sliding_context_len = self.max_tokens - 1
        B, gt_len = context.shape[:2]
        if sliding_context_len < gt_len:
            raise ValueError("sliding_context_len is expected to be >= length of initial context,"
                             f"got {sliding_context_len}. If you are trying to use max context, "
                             "consider specifying sliding_context_len=-1.")
        chunk = self.chunk_size if self.use_causal_mask else self.max_tokens
        xs, cur, record = context, gt_len, None
        while cur < length:
            if record is not None:
                raise ValueError("return_all is not supported if using sliding window.")
            c = min(sliding_context_len, cur)
            h = min(length - cur, self.max_tokens - c)
            h = min(h, chunk) if chunk > 0 else h
            window = torch.cat([xs[:, -c:], torch.zeros((B, h, *self.x_shape), dtype=xs.dtype, device=xs.device)], 1)
            n_generated = cur - max(cur - c, gt_len)
            mask = np.ones((B, c + h), dtype=np.int64)            # 1 = ground-truth context
            if n_generated > 0:
                mask[:, c - n_generated:c] = 2                     # 2 = generated context
            mask[:, c:] = 0                                        # 0 = to be generated
            cond = None
            if conditions is not None:
                cond_len = c + h if self.use_causal_mask else self.max_tokens
                if self.external_cond_type == "label":
                    cond = conditions
                elif self.external_cond_type == "action":
                    cond = conditions[:, cur - c: cur - c + cond_len]
                else:
                    raise ValueError(f"Unknown external condition type: {self.external_cond_type}. "
                                     "Supported types are 'label' and 'action'.")
            new, record = self._window_sampler()(B, length=c + h, context=window, context_mask=torch.from_numpy(mask),
                                                 conditions=cond, guidance_fn=guidance_fn,
                                                 reconstruction_guidance=reconstruction_guidance,
                                                 history_guidance=history_guidance, return_all=return_all)
            xs = torch.cat([xs, new[:, -h:]], 1)
            cur = xs.shape[1]
        return xs, record

    def _window_conditions(self, conditions: Tensor, nfe: int):
        """Conditioning of all branch rows `(b h g)` of one window (rows of a sample share its conditions)."""
        return self._process_conditions(conditions.repeat_interleave(nfe, dim=0).clone(), None)

    def _model_in_buffer(self, rows: int, T: int, dev) -> Tensor:
        """Branch-input tensor the fused sampler kernel writes into: the backbone's static (graph-captured) input
        when it offers one, so no copy sits between K4 and the forward."""
        model = self.diffusion_model.model
        bg = self.mesh.branch_group if self.mesh is not None else None
        if hasattr(model, "input_buffer") and bg is None:
            return model.input_buffer(rows, T, self.model_in_dtype, dev)
        return torch.empty((rows, T, *self.x_shape), dtype=self.model_in_dtype, device=dev)

    # ------------------------------------------------------------------ multi-GPU
    def _backbone_rows(self, model_in, levels, cond, cond_mask, B: int, nfe: int):
        """Backbone forward over the (b, j) branch rows; with a branch group each member runs its share of the
        branches and the outputs are all-gathered (the only per-step collective of the path)."""
        dm = self.diffusion_model
        bg = self.mesh.branch_group if self.mesh is not None else None
        if bg is None or nfe == 1:
            return dm.model(model_in, levels, cond, cond_mask, out_dtype=torch.float32)
        from dfot_b200 import distributed as D
        rows = torch.tensor(D.branch_rows(B, nfe, bg), device=model_in.device)
        sel = lambda t: None if t is None else t.index_select(0, rows)
        local = dm.model(sel(model_in), sel(levels), sel(cond), sel(cond_mask), out_dtype=torch.float32)
        return D.gather_branch_outputs(local, B, nfe, bg)

    @torch.no_grad()
    def sample_sharded(self, xs: Tensor, conditions: Optional[Tensor] = None,
                       n_context_tokens: Optional[int] = None) -> Tensor:
        """`_predict_videos` over the dp x br mesh: this rank's sample shard is rolled out (branch rows split inside
        its branch group) and the finished samples of all shards are all-gathered; every rank returns the full batch.
        xs / conditions hold the FULL batch on every rank (synthetic or broadcast by the caller)."""
        from dfot_b200 import distributed as D
        mesh = self.mesh
        n_ctx = n_context_tokens if n_context_tokens is not None else self.n_context_tokens
        if mesh is None:
            return self._predict_videos(xs, n_ctx, conditions)
        if xs.shape[0] < mesh.dp:
            # fewer samples than dp shards (the 200-frame single-sample rollout of BASELINE config[3]): every shard rolls
            # the whole batch — keyframe windows replicated (branches still split inside the branch group), interpolation
            # chunk batches spread over the dp axis.  All ranks must share the noise seed.
            self.shard_chunks = True
            try:
                return self._predict_videos(xs, n_ctx, conditions)
            finally:
                self.shard_chunks = False
        counts = [len(range(*D.shard_batch(xs.shape[0], mesh.dp, d).indices(xs.shape[0]))) for d in range(mesh.dp)]
        sl = D.shard_batch(xs.shape[0], mesh.dp, mesh.dp_index)
        # Every rank is seeded identically (the replicated paths above and inside a branch group need ONE noise stream);
        # the shards of a sample-sharded batch must not share their noise, so each dp shard draws from its own generator,
        # derived from the common seed (members of a branch group share dp_index, hence the stream).
        dm = self.diffusion_model
        prev = dm.generator
        if mesh.dp > 1 and dm.noise_source is None and prev is None:
            dm.generator = self._shard_generator(xs.device, mesh.dp_index)
        try:
            local = self._predict_videos(xs[sl].contiguous(), n_ctx, None if conditions is None else conditions[sl])
        finally:
            dm.generator = prev
        return D.gather_samples(local.contiguous(), mesh, counts)

    def _shard_generator(self, device, dp_index: int) -> torch.Generator:
        """Noise stream of dp shard `dp_index`: seeded once from the process-wide seed (`torch.manual_seed`, identical on
        all ranks) and then advanced batch after batch; re-seeding the process starts a new stream."""
        key = (str(device), torch.initial_seed(), dp_index)
        if getattr(self, "_shard_gen_key", None) != key:
            g = torch.Generator(device=device)
            g.manual_seed((torch.initial_seed() + 0x9E3779B1 * (dp_index + 1)) % (1 << 63))
            self._shard_gen, self._shard_gen_key = g, key
        return self._shard_gen

    # ------------------------------------------------------------------ window planning (host only)
    def plan_window(self, mask: np.ndarray, horizon: int, padding: int,
                    history_guidance: HistoryGuidance) -> List[sp.StepPlan]:
        """Everything integer / per-frame-scalar about one window, for all of its steps.
        mask: int64 [B, horizon] (already padded with -1).  Pure host code (numpy), no device access."""
        B = mask.shape[0]
        dm = self.diffusion_model
        # scheduling matrix: repeat over batch, context → -1, drop leading duplicate rows (:642-657)
        S = self._generate_scheduling_matrix(horizon - padding, padding).numpy()
        S = np.repeat(S[:, None, :], B, axis=1)
        if not self.is_full_sequence:
            S = np.where(mask[None] >= 1, -1, S)
        changed = (S[1:] != S[:-1]).reshape(S.shape[0] - 1, -1).any(axis=1)
        S = S[int(np.argmax(changed)):]
        plans: List[sp.StepPlan] = []
        for m in range(S.shape[0] - 1):
            frm, to = S[m], S[m + 1]
            mask = np.where((mask == 0) & (frm == -1), 2, mask)           # (:675-679)
            plans.append(history_guidance.plan_step(dm.host_tables, mask, frm, to, self.is_full_sequence,
                                                    dm.is_continuous, dm.precond_scale))
        return plans

    # ------------------------------------------------------------------ refinement sampling (fork-only, :765-1008)
    def _window_sampler(self) -> Callable:
        rs = self.cfg.get("refinement_sampling")
        return self._sample_sequence_refine if rs is not None and rs.get("enabled") else self._sample_sequence

    @torch.no_grad()
    def _sample_sequence_refine(self, batch_size: int, goback_length: Optional[int] = None, n_goback: Optional[int] = None,
                                length: Optional[int] = None, context: Optional[Tensor] = None,
                                context_mask: Optional[Tensor] = None, conditions: Optional[Tensor] = None,
                                guidance_fn: Optional[Callable] = None, reconstruction_guidance: float = 0.0,
                                history_guidance: Optional[HistoryGuidance] = None, return_all: bool = False,
                                pbar=None, dry_run: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
        """The window sampler over the refinement walk (`_generate_refine_scheduling_matrix`): a row whose LAST column's
        level decreases is a denoising step (backbone + fused DDIM update), any other row re-noises every token from its
        level up to the next one (`q_sample_from_x_k`, one K4 launch).  As in the reference, the loop only works with a
        single history-guidance branch (its `q_sample(context, to_noise_levels)` at :983 mixes (B, T) data with
        (B·nfe, T) levels) — that draw is kept for the RNG stream, its result is discarded there as well (:986-989).
        With the discrete cosine schedule ᾱ[-1] == 0, so the re-noising scale of context tokens (level -1) is 0/0 and
        the reference's rollout is NaN (quirk Q11); the arithmetic here is the same."""
        rs = self.cfg.refinement_sampling
        goback_length = rs.goback_length if goback_length is None else goback_length
        n_goback = rs.n_goback if n_goback is None else n_goback
        x_shape = self.x_shape
        if guidance_fn is not None or reconstruction_guidance > 0:
            raise NotImplementedError("guidance_fn / reconstruction guidance needs autograd through the backbone "
                                      "and is outside the dfot_b200 scope (SURVEY.md §3.4)")
        if dry_run:
            raise NotImplementedError("refinement sampling is not sharded over interpolation chunk batches")
        if context is None:
            raise ValueError("context must be provided")
        if length is None:
            length = context.shape[1]
        if length > self.max_tokens:
            raise ValueError(f"length is expected to <={self.max_tokens}, got {length}.")
        if context_mask is None:
            raise ValueError("context_mask must be provided if context is given.")
        if context.shape[0] != batch_size:
            raise ValueError(f"context batch size is expected to be {batch_size} but got {context.shape[0]}.")
        if context.shape[1] != length:
            raise ValueError(f"context length is expected to be {length} but got {context.shape[1]}.")
        if tuple(context.shape[2:]) != tuple(x_shape):
            raise ValueError(f"context shape not compatible with x_stacked_shape {x_shape}.")
        if tuple(context.shape[:2]) != tuple(context_mask.shape):
            raise ValueError("context and context_mask must have the same shape.")
        dev, dm, B = context.device, self.diffusion_model, batch_size
        horizon = length if self.use_causal_mask else self.max_tokens
        padding = horizon - length
        if dm.noise_source is None and self.generator is not None:
            x = torch.randn((B, horizon, *x_shape), device=dev, generator=self.generator)
        else:
            x = dm.randn((B, horizon, *x_shape), dev)
        x = torch.clamp(x, -self.clip_noise, self.clip_noise)
        mask = context_mask.detach().cpu().numpy().astype(np.int64)
        if padding > 0:
            context = torch.cat([context, torch.zeros((B, padding, *x_shape), dtype=context.dtype, device=dev)], 1)
            mask = np.concatenate([mask, -np.ones((B, padding), dtype=np.int64)], 1)
        if history_guidance is None:
            history_guidance = HistoryGuidance.conditional(timesteps=self.timesteps)
        x = torch.where(self._extend_x_dim(torch.from_numpy(mask).to(dev)) >= 1, context.float(), x).contiguous()
        S = self._generate_refine_scheduling_matrix(horizon - padding, goback_length, n_goback, padding).numpy()
        S = np.repeat(S[:, None, :], B, axis=1)
        S = np.where(mask[None] >= 1, -1, S)                                # (:889-891)
        T = horizon
        cond = None if conditions is None else self._window_conditions(conditions.to(dev), 1)
        record = [] if return_all else None
        spare = torch.empty_like(x)
        for m in range(S.shape[0] - 1):
            frm, to = S[m], S[m + 1]
            if frm[0, -1] > to[0, -1]:
                mask = np.where((mask == 0) & (frm == -1), 2, mask)
                if return_all:
                    record.append(x.clone())
                p = history_guidance.plan_step(dm.host_tables, mask, frm, to, self.is_full_sequence, dm.is_continuous,
                                               dm.precond_scale)
                if p.nfe != 1:
                    raise NotImplementedError("refinement sampling runs with one history-guidance branch only (the "
                                              "reference's loop fails for nfe > 1, dfot_video.py:983)")
                # RNG: q_sample noise of the history tokens, then (full manager) the excluded-token noise — always drawn
                nh = dm.clipped_noise((p.n_hist_rows, T, *x_shape), dev) if p.n_hist_rows else None
                ne = dm.randn((B, T, *x_shape), dev) if p.draws_excluded_noise else None
                model_in = self._model_in_buffer(B, T, dev)
                ops.sampler_step_hg(x, None, model_in, None, sp.to_device_bytes(p.prepare, dev), None, nh, ne, B, 1, T)
                self.nfe_rows_planned += B
                lv = torch.from_numpy(p.levels).to(dev, non_blocking=True)
                cm = None if p.cond_mask is None else torch.from_numpy(p.cond_mask).to(dev, non_blocking=True)
                out = self._backbone_rows(model_in, lv, cond, cm, B, 1)
                self.nfe_rows += B
                nd = dm.clipped_noise((B, T, *x_shape), dev)                 # RNG: DDIM noise, drawn even when eta == 0
                trace_in = model_in.float().clone() if self.trace is not None else None
                ops.sampler_step_hg(x, out, None, sp.to_device_bytes(p.update, dev), None,
                                    nd if dm.host_tables.uses_step_noise else None, None, None, B, 1, T)
                dm.clipped_noise((B, T, *x_shape), dev)                      # RNG: q_sample(context, to) of :983, unused
                if self.trace is not None:
                    self.trace.append(dict(model_in=trace_in, levels_from=p.levels_from, levels_to=p.levels_to,
                                           cond_mask=p.cond_mask, model_out=out.float().clone(), x_after=x.clone(),
                                           context_mask=p.context_mask))
            else:
                noise = dm.clipped_noise((B, T, *x_shape), dev)
                prep = dm.renoise_table(frm, to)
                ops.sampler_step_hg(x, None, spare, None, sp.to_device_bytes(prep, dev), None, noise, None, B, 1, T)
                x, spare = spare, x
        if return_all:
            record.append(x.clone())
            record = torch.stack(record)
        if padding > 0:
            x = x[:, :-padding]
            record = record[:, :, :-padding] if return_all else None
        return x, record

    # ------------------------------------------------------------------ the hot loop
    @torch.no_grad()
    def _sample_sequence(self, batch_size: int, length: Optional[int] = None, context: Optional[Tensor] = None,
                         context_mask: Optional[Tensor] = None, conditions: Optional[Tensor] = None,
                         guidance_fn: Optional[Callable] = None, reconstruction_guidance: float = 0.0,
                         history_guidance: Optional[HistoryGuidance] = None, return_all: bool = False,
                         pbar=None, dry_run: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
        """The window sampler (reference `_sample_sequence`, dfot_video.py:530-763).  `dry_run` (multi-GPU chunk
        sharding) draws exactly the noise a real call would — so every rank's generator stays where the single-GPU run
        would have it — but launches no kernel; the returned tensor is a placeholder for the owner's broadcast."""
        x_shape = self.x_shape
        if guidance_fn is not None or reconstruction_guidance > 0:
            raise NotImplementedError("guidance_fn / reconstruction guidance needs autograd through the backbone "
                                      "and is outside the dfot_b200 scope (SURVEY.md §3.4)")
        if context is None:
            # the reference's context=None branch is broken (torch.zeros_like(tuple), :616-619) and unreachable
            raise ValueError("context must be provided")
        if length is None:
            length = context.shape[1]
        if length > self.max_tokens:
            raise ValueError(f"length is expected to <={self.max_tokens}, got {length}.")
        if context_mask is None:
            raise ValueError("context_mask must be provided if context is given.")
        if context.shape[0] != batch_size:
            raise ValueError(f"context batch size is expected to be {batch_size} but got {context.shape[0]}.")
        if context.shape[1] != length:
            raise ValueError(f"context length is expected to be {length} but got {context.shape[1]}.")
        if tuple(context.shape[2:]) != tuple(x_shape):
            raise ValueError(f"context shape not compatible with x_stacked_shape {x_shape}.")
        if tuple(context.shape[:2]) != tuple(context_mask.shape):
            raise ValueError("context and context_mask must have the same shape.")
        dev = context.device
        dm = self.diffusion_model
        B = batch_size
        horizon = length if self.use_causal_mask else self.max_tokens
        padding = horizon - length

        # ---- RNG ①: initial noise (:607-612)
        if dm.noise_source is None and self.generator is not None:
            x = torch.randn((B, horizon, *x_shape), device=dev, generator=self.generator)
        else:
            x = dm.randn((B, horizon, *x_shape), dev)
        x = torch.clamp(x, -self.clip_noise, self.clip_noise)

        mask = context_mask.detach().cpu().numpy().astype(np.int64)
        if padding > 0:   # -1 marks padding frames (:620-630); they carry noise at level T-1 and ARE attended (Q4)
            context = torch.cat([context, torch.zeros((B, padding, *x_shape), dtype=context.dtype, device=dev)], 1)
            mask = np.concatenate([mask, -np.ones((B, padding), dtype=np.int64)], 1)
        if history_guidance is None:
            history_guidance = HistoryGuidance.conditional(timesteps=self.timesteps)
        mask_dev = torch.from_numpy(mask).to(dev)
        x = torch.where(self._extend_x_dim(mask_dev) >= 1, context.float(), x).contiguous()

        plans = self.plan_window(mask, horizon, padding, history_guidance)
        n_steps = len(plans)
        tb = dm.host_tables
        upd_dev = [sp.to_device_bytes(p.update, dev) for p in plans]
        prep_dev = [sp.to_device_bytes(p.prepare, dev) for p in plans]
        lvl_dev = [torch.from_numpy(p.levels).to(dev, non_blocking=True) for p in plans]
        cm_dev = [None if p.cond_mask is None else torch.from_numpy(p.cond_mask).to(dev, non_blocking=True)
                  for p in plans]
        cond_cache: Dict[int, Tensor] = {}

        def cond_for(nfe: int):
            if conditions is None or dry_run:
                return None
            if nfe not in cond_cache:   # constant over the window (the reference recomputes it every step, :732-743)
                cond_cache[nfe] = self._window_conditions(conditions.to(dev), nfe)
            return cond_cache[nfe]

        def draw_prepare_noise(p: sp.StepPlan):
            # RNG ②: q_sample noise, then (full manager only) the excluded-token noise — always drawn
            nh = dm.clipped_noise((p.n_hist_rows, horizon, *x_shape), dev) if p.n_hist_rows else None
            ne = dm.randn((B * p.nfe, horizon, *x_shape), dev) if p.draws_excluded_noise else None
            return nh, ne

        record = [] if return_all else None
        T = horizon
        model_in = None
        def k4(*args, hist_rows: int = 0):
            if not dry_run:
                ops.sampler_step_hg(*args, max_noise_row=hist_rows - 1 if hist_rows else None)

        for m, p in enumerate(plans):
            if return_all:
                record.append(x.clone())
            if m == 0:
                nh, ne = draw_prepare_noise(p)
                model_in = self._model_in_buffer(B * p.nfe, T, dev)
                k4(x, None, model_in, None, prep_dev[0], None, nh, ne, B, p.nfe, T, hist_rows=p.n_hist_rows)
            out = None
            self.nfe_rows_planned += B * p.nfe
            if not dry_run:
                out = self._backbone_rows(model_in, lvl_dev[m], cond_for(p.nfe), cm_dev[m], B, p.nfe)
                self.nfe_rows += B * p.nfe
            # RNG ③: the step's noise (DDIM sigma / DDPM) — drawn even when eta == 0, to stay aligned with the reference's stream
            nd = dm.clipped_noise((B * p.nfe, T, *x_shape), dev)
            nd = nd if tb.uses_step_noise else None
            trace_in = model_in.float().clone() if self.trace is not None else None
            nxt = plans[m + 1] if m + 1 < n_steps else None
            if nxt is None:
                k4(x, out, None, upd_dev[m], None, nd, None, None, B, p.nfe, T)
            else:
                nh, ne = draw_prepare_noise(nxt)
                nxt_in = model_in if nxt.nfe == p.nfe else self._model_in_buffer(B * nxt.nfe, T, dev)
                if nxt.nfe == p.nfe:   # one fused launch: update + combine + revert + next-step prepare
                    k4(x, out, nxt_in, upd_dev[m], prep_dev[m + 1], nd, nh, ne, B, p.nfe, T, hist_rows=nxt.n_hist_rows)
                else:                  # branch count changes between steps: split into update and prepare launches
                    k4(x, out, None, upd_dev[m], None, nd, None, None, B, p.nfe, T)
                    k4(x, None, nxt_in, None, prep_dev[m + 1], None, nh, ne, B, nxt.nfe, T, hist_rows=nxt.n_hist_rows)
                model_in = nxt_in
            if self.trace is not None and not dry_run:
                self.trace.append(dict(model_in=trace_in, levels_from=p.levels_from, levels_to=p.levels_to,
                                       cond_mask=p.cond_mask, model_out=out.float().clone(), x_after=x.clone(),
                                       context_mask=p.context_mask))
        if return_all:
            record.append(x.clone())
            record = torch.stack(record)
        if padding > 0:
            x = x[:, :-padding]
            record = record[:, :, :-padding] if return_all else None
        return x, record
