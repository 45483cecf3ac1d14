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
        self.row_shard = None                   # distributed.RowShard: forward rows dealt over the whole world (sample_sharded)
        self._row_cond_cache: Dict = {}

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
            # every chunk is processed (the reference's conditions=None path drops the last partial batch — quirk Q10,
            # not replicated).  The chunk batches of one round are independent (SURVEY.md §8e axis 3): on one GPU they are
            # sampled one after the other like the reference does; with a row shard they advance in lockstep, so that the
            # rows of the WHOLE round are what the GPUs share (`_run_lockstep`; same noise per batch either way).
            spans = [(s, min(rows, s + mb)) for s in range(0, rows, mb)]

            def maker(s, e):
                return lambda dry: _WindowRun(self, e - s, None, ctx[s:e], torch.from_numpy(msk[s:e].astype(np.int64)),
                                              None if cnd is None else cnd[s:e], None, 0.0, guidance, False, dry=dry)

            sampler = self._window_sampler()
            if self.row_shard is not None and sampler == self._sample_sequence and self.diffusion_model.noise_can_fork():
                outs = [o for o, _ in self._run_lockstep([maker(s, e) for s, e in spans])]
            else:
                outs = [sampler(batch_size=e - s, context=ctx[s:e],
                                context_mask=torch.from_numpy(msk[s:e].astype(np.int64)),
                                conditions=None if cnd is None else cnd[s:e], history_guidance=guidance)[0]
                        for s, e in spans]
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

    def _conditions_follow_levels(self) -> bool:
        """True when `_process_conditions` reads the step's noise levels (pose conditioning under `temporal` history
        guidance); the window then caches one conditioning per distinct mask of top-level frames instead of one."""
        return False

    def _window_conditions(self, conditions: Tensor, nfe: int, levels_from=None):
        """Conditioning of all branch rows `(b h g)` of one window (rows of a sample share its conditions)."""
        return self._process_conditions(conditions.repeat_interleave(nfe, dim=0).clone(), None)

    def _model_in_buffer(self, rows: int, T: int, dev) -> Tensor:
        """Branch-input tensor the fused sampler kernel writes into: the backbone's static (graph-captured) input
        when it offers one, so no copy sits between K4 and the forward."""
        model = self.diffusion_model.model
        if hasattr(model, "input_buffer") and self._active_row_shard() is None:
            return model.input_buffer(rows, T, self.model_in_dtype, dev)
        return torch.empty((rows, T, *self.x_shape), dtype=self.model_in_dtype, device=dev)

    # ------------------------------------------------------------------ multi-GPU
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
            # fewer samples than dp shards (the 200-frame single-sample rollout of BASELINE config[3]): the sampler state
            # is replicated on every rank — all ranks must share the noise seed — and the backbone forward rows (keyframe
            # windows; whole interpolation rounds in lockstep) are dealt over ALL ranks.  Every rank ends with the video.
            self.row_shard = D.RowShard()
            try:
                return self._predict_videos(xs, n_ctx, conditions)
            finally:
                self.row_shard = None
                self._row_cond_cache = {}
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
                out = dm.model(model_in, lv, cond, cm, out_dtype=torch.float32)
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
        """The window sampler (reference `_sample_sequence`, dfot_video.py:530-763): plan the window, then per step one
        backbone forward over the branch rows and one fused K4 launch.  `dry_run` draws exactly the noise a real call
        would — so the generator ends where the real call leaves it — but launches no kernel."""
        run = _WindowRun(self, batch_size, length, context, context_mask, conditions, guidance_fn, reconstruction_guidance,
                         history_guidance, return_all, dry=dry_run)
        run.begin()
        while not run.done:
            run.advance(None if dry_run else self._forward_runs([run])[0])
        return run.result()

    # ------------------------------------------------------------------ backbone forward over the rows of >= 1 windows
    def _active_row_shard(self):
        """How the forward rows are spread over GPUs: the whole world (`row_shard`, replicated sampler state), the members
        of this sample shard's branch group (mesh br > 1), or not at all."""
        if self.row_shard is not None:
            return self.row_shard
        bg = self.mesh.branch_group if self.mesh is not None else None
        if bg is None:
            return None
        if getattr(self, "_bg_shard", None) is None or self._bg_shard.group is not bg.group:
            from dfot_b200 import distributed as D
            self._bg_shard = D.RowShard(world=bg.size, rank=bg.rank, group=bg.group)
        return self._bg_shard

    def _forward_runs(self, runs: List["_WindowRun"]) -> List[Tensor]:
        """ONE backbone forward for the current step of every window in `runs` (their branch rows concatenated) — on a
        single GPU over all rows; with a row shard over this rank's contiguous block of the rows, followed by the only
        per-step collective of the path: an all_gather_into_tensor of the outputs."""
        dm = self.diffusion_model
        ins = [r.step_inputs() for r in runs]
        counts = [i["model_in"].shape[0] for i in ins]
        n = sum(counts)
        self.nfe_rows_planned += n
        rs = self._active_row_shard()
        if rs is None and len(runs) == 1:
            i = ins[0]
            self.nfe_rows += n
            return [dm.model(i["model_in"], i["levels"], i["cond"], i["cond_mask"], out_dtype=torch.float32)]
        start, stop = (0, n) if rs is None else rs.block(n)[1:]
        pieces, off = [], 0
        for k, c in enumerate(counts):          # (window, first row, last row) of the rows in [start, stop)
            lo, hi = max(start, off), min(stop, off + c)
            if lo < hi:
                pieces.append((k, lo - off, hi - off))
            off += c
        local = None
        if pieces:
            cat = lambda key: torch.cat([ins[k][key][lo:hi] for k, lo, hi in pieces], 0)
            cond = self._gather_conditions([(runs[k], ins[k]["cond"], lo, hi) for k, lo, hi in pieces])
            cm = None if ins[pieces[0][0]]["cond_mask"] is None else cat("cond_mask")
            local = dm.model(cat("model_in"), cat("levels"), cond, cm, out_dtype=torch.float32)
            self.nfe_rows += stop - start
        if rs is None:
            full = local
        else:
            ref = ins[0]["model_in"]
            full = rs.gather(local, n, tuple(ref.shape[1:]), torch.float32, ref.device)
        return list(full.split(counts, 0))

    def _gather_conditions(self, pieces):
        """Conditioning of the local rows of a multi-window forward.  Tensors (actions, labels) are concatenated; the
        camera-pose handles of the U-ViT are merged into ONE handle over the windows this rank actually forwards, cached per
        set of pieces so that the pose modulation cache is filled once per round, not once per step."""
        conds = [c for _, c, _, _ in pieces]
        if conds[0] is None:
            return None
        if torch.is_tensor(conds[0]):
            return torch.cat([c[lo:hi] for _, c, lo, hi in pieces], 0)
        # the cache entry HOLDS the handles it was built from (identity comparison; ids of dead objects get re-used)
        held = self._row_cond_cache.get("pieces")
        if held is not None and len(held) == len(pieces) and all(
                h[0] is c and h[1:] == (lo, hi) for h, (_, c, lo, hi) in zip(held, pieces)):
            return self._row_cond_cache["merged"]
        from .backbones.u_vit.u_vit3d_pose import PoseCondition
        cams, row_map = [], []
        for _, c, lo, hi in pieces:
            used = sorted(set(c.row_map[lo:hi]))
            base = sum(x.shape[0] for x in cams)
            cams.append(c.cams[used])
            row_map += [base + used.index(j) for j in c.row_map[lo:hi]]
        merged = PoseCondition(torch.cat(cams, 0), row_map)
        self._row_cond_cache = {"pieces": [(c, lo, hi) for _, c, lo, hi in pieces], "merged": merged}
        return merged

    @torch.no_grad()
    def _run_lockstep(self, makers: List[Callable]) -> List[Tuple[Tensor, Optional[Tensor]]]:
        """Independent windows that the reference samples one after the other (the chunk batches of an interpolation round,
        dfot_video.py:284-358), advanced TOGETHER: one backbone forward per step over the rows of all of them — which is
        what lets 8 GPUs share a round of 70 rows evenly instead of 4 chunks at a time.  Every window still draws exactly
        the noise the sequential order gives it: the stream position at the start of each window is found by replaying
        the preceding windows' draws (no kernels), and each window then draws from its own position.
        makers[i](dry) -> _WindowRun."""
        dm = self.diffusion_model
        dev = self.device
        states = []
        for mk in makers:
            states.append(dm.noise_get_state(dev))
            r = mk(True)
            r.begin()
            while not r.done:
                r.advance(None)
        final = dm.noise_get_state(dev)
        runs = []
        for mk, st in zip(makers, states):
            dm.noise_set_state(dev, st)
            r = mk(False)
            r.begin()
            r.noise_state = dm.noise_get_state(dev)
            runs.append(r)
        while True:
            active = [r for r in runs if not r.done]
            if not active:
                break
            for r, out in zip(active, self._forward_runs(active)):
                dm.noise_set_state(dev, r.noise_state)
                r.advance(out)
                r.noise_state = dm.noise_get_state(dev)
        dm.noise_set_state(dev, final)
        return [r.result() for r in runs]


class _WindowRun:
    """One denoising window as a resumable state machine: `begin()` builds the first step's branch inputs, `step_inputs()`
    hands them to a backbone forward, `advance(out)` applies the fused K4 step (DDIM / DDPM update, guidance combine,
    context revert, next step's branch inputs).  RNG order = the reference's: (1) initial noise; per step (2) history
    re-noising + excluded-token noise of the step's branch inputs, (3) the step's update noise."""

    def __init__(self, algo: DFoTVideo, batch_size: int, length: Optional[int], context: Optional[Tensor],
                 context_mask: Optional[Tensor], conditions: Optional[Tensor], guidance_fn: Optional[Callable],
                 reconstruction_guidance: float, history_guidance: Optional[HistoryGuidance], return_all: bool,
                 dry: bool = False):
        x_shape = algo.x_shape
        if guidance_fn is not None or reconstruction_guidance > 0:
            raise NotImplementedError("guidance_fn / reconstruction guidance needs autograd through the backbone "
                                      "and is outside the dfot_b200 scope (SURVEY.md §3.4)")
        if context is None:
            # the reference's context=None branch is broken (torch.zeros_like(tuple), :616-619) and unreachable
            raise ValueError("context must be provided")
        if length is None:
            length = context.shape[1]
        if length > algo.max_tokens:
            raise ValueError(f"length is expected to <={algo.max_tokens}, got {length}.")
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
        self.algo, self.dry, self.return_all = algo, dry, return_all
        self.dev = dev = context.device
        self.dm = dm = algo.diffusion_model
        self.B = B = batch_size
        horizon = length if algo.use_causal_mask else algo.max_tokens
        self.T, self.padding = horizon, horizon - length
        self.conditions = conditions
        self.noise_state = None

        # ---- RNG (1): initial noise (:607-612)
        if dry and not (dm.noise_source is None and algo.generator is not None):
            dm.skip_randn((B, horizon, *x_shape), dev)      # a dry run wants the stream position, not the values
            x = torch.zeros((B, horizon, *x_shape), device=dev)
        elif dm.noise_source is None and algo.generator is not None:
            x = torch.randn((B, horizon, *x_shape), device=dev, generator=algo.generator)
        else:
            x = dm.randn((B, horizon, *x_shape), dev)
        x = torch.clamp(x, -algo.clip_noise, algo.clip_noise)
        mask = context_mask.detach().cpu().numpy().astype(np.int64)
        if self.padding > 0:   # -1 marks padding frames (:620-630); they carry noise at level T-1 and ARE attended (Q4)
            context = torch.cat([context, torch.zeros((B, self.padding, *x_shape), dtype=context.dtype, device=dev)], 1)
            mask = np.concatenate([mask, -np.ones((B, self.padding), dtype=np.int64)], 1)
        if history_guidance is None:
            history_guidance = HistoryGuidance.conditional(timesteps=algo.timesteps)
        mask_dev = torch.from_numpy(mask).to(dev)
        self.x = torch.where(algo._extend_x_dim(mask_dev) >= 1, context.float(), x).contiguous()

        self.plans = algo.plan_window(mask, horizon, self.padding, history_guidance)
        self.n_steps = len(self.plans)
        if not dry:
            self.upd_dev = [sp.to_device_bytes(p.update, dev) for p in self.plans]
            self.prep_dev = [sp.to_device_bytes(p.prepare, dev) for p in self.plans]
            self.lvl_dev = [torch.from_numpy(p.levels).to(dev, non_blocking=True) for p in self.plans]
            self.cm_dev = [None if p.cond_mask is None else torch.from_numpy(p.cond_mask).to(dev, non_blocking=True)
                           for p in self.plans]
        self.cond_cache: Dict[object, Tensor] = {}
        self.record = [] if return_all else None
        self.model_in = None
        self.m = 0

    @property
    def done(self) -> bool:
        return self.m >= self.n_steps

    def _cond_for(self, p: sp.StepPlan):
        if self.conditions is None or self.dry:
            return None
        nfe, key, levels = p.nfe, p.nfe, None
        if self.algo._conditions_follow_levels():   # `temporal` HG + poses: depends on which frames sit at the top level
            levels = p.levels_from
            key = (nfe, (levels == self.algo.timesteps - 1).tobytes())
        if key not in self.cond_cache:   # constant over the window (the reference recomputes it every step, :732-743)
            self.cond_cache[key] = self.algo._window_conditions(self.conditions.to(self.dev), nfe, levels_from=levels)
        return self.cond_cache[key]

    def _draw_prepare_noise(self, p: sp.StepPlan):
        # RNG (2): q_sample noise, then (full manager only) the excluded-token noise — always drawn
        shape = (self.T, *self.algo.x_shape)
        if self.dry:
            if p.n_hist_rows:
                self.dm.skip_randn((p.n_hist_rows, *shape), self.dev)
            if p.draws_excluded_noise:
                self.dm.skip_randn((self.B * p.nfe, *shape), self.dev)
            return None, None
        nh = self.dm.clipped_noise((p.n_hist_rows, *shape), self.dev) if p.n_hist_rows else None
        ne = self.dm.randn((self.B * p.nfe, *shape), self.dev) if p.draws_excluded_noise else None
        return nh, ne

    def _k4(self, *args, hist_rows: int = 0):
        if not self.dry:
            ops.sampler_step_hg(*args, max_noise_row=hist_rows - 1 if hist_rows else None)

    def begin(self) -> None:
        if self.n_steps == 0:
            return
        p = self.plans[0]
        nh, ne = self._draw_prepare_noise(p)
        if not self.dry:
            self.model_in = self.algo._model_in_buffer(self.B * p.nfe, self.T, self.dev)
        self._k4(self.x, None, self.model_in, None, None if self.dry else self.prep_dev[0], None, nh, ne, self.B, p.nfe,
                 self.T, hist_rows=p.n_hist_rows)

    def step_inputs(self) -> dict:
        p = self.plans[self.m]
        return dict(model_in=self.model_in, levels=self.lvl_dev[self.m], cond=self._cond_for(p),
                    cond_mask=self.cm_dev[self.m])

    def advance(self, out: Optional[Tensor]) -> None:
        algo, dm, m, B, T = self.algo, self.dm, self.m, self.B, self.T
        p = self.plans[m]
        if self.return_all:
            self.record.append(self.x.clone())
        # RNG (3): the step's noise (DDIM sigma / DDPM) — drawn even when eta == 0, to stay aligned with the reference
        nd = None
        if dm.host_tables.uses_step_noise and not self.dry:
            nd = dm.clipped_noise((B * p.nfe, T, *algo.x_shape), self.dev)
        else:               # sigma == 0 everywhere (or a dry run): advance the stream, skip the values
            dm.skip_randn((B * p.nfe, T, *algo.x_shape), self.dev)
        tracing = algo.trace is not None and not self.dry
        trace_in = self.model_in.float().clone() if tracing else None
        nxt = self.plans[m + 1] if m + 1 < self.n_steps else None
        upd = None if self.dry else self.upd_dev[m]
        if nxt is None:
            self._k4(self.x, out, None, upd, None, nd, None, None, B, p.nfe, T)
        else:
            nh, ne = self._draw_prepare_noise(nxt)
            prep = None if self.dry else self.prep_dev[m + 1]
            if nxt.nfe == p.nfe:   # one fused launch: update + combine + revert + next-step prepare
                self._k4(self.x, out, self.model_in, upd, prep, nd, nh, ne, B, p.nfe, T, hist_rows=nxt.n_hist_rows)
            else:                  # branch count changes between steps: split into update and prepare launches
                self._k4(self.x, out, None, upd, None, nd, None, None, B, p.nfe, T)
                if not self.dry:
                    self.model_in = algo._model_in_buffer(B * nxt.nfe, T, self.dev)
                self._k4(self.x, None, self.model_in, None, prep, None, nh, ne, B, nxt.nfe, T, hist_rows=nxt.n_hist_rows)
        if tracing:
            algo.trace.append(dict(model_in=trace_in, levels_from=p.levels_from, levels_to=p.levels_to,
                                   cond_mask=p.cond_mask, model_out=out.float().clone(), x_after=self.x.clone(),
                                   context_mask=p.context_mask))
        self.m += 1

    def result(self) -> Tuple[Tensor, Optional[Tensor]]:
        x, record = self.x, None
        if self.return_all:
            self.record.append(x.clone())
            record = torch.stack(self.record)
        if self.padding > 0:
            x = x[:, :-self.padding]
            record = record[:, :, :-self.padding] if self.return_all else None
        return x, record
