"""ORACLE (test infrastructure): the scheduling-matrix sampler driver.

Restates algorithms/dfot/dfot_video.py
  _predict_videos :114-179, _interpolate_videos :181-360 (plan :219-261),
  _predict_sequence :362-514, _sample_sequence :516-763
and algorithms/common/base_pytorch_video_algo.py
  _process_conditions :635-664, _pad_to_max_tokens :666-682, token maths :986-1033.
Noise is drawn in the reference's order (SURVEY.md §8a "RNG contract") from the
injected ``randn`` / ``randn_like`` callables (default: torch global generator).
"""
from typing import Callable, List, Optional

import torch

from . import history_guidance as hg
from .diffusion import Diffusion
from .schedule import refine_scheduling_matrix, scheduling_matrix


def interpolation_plan(known: torch.Tensor, max_tokens: int) -> List[List[torch.Tensor]]:
    """dfot_video.py:219-261. known: bool [T]; returns rounds of frame-index chunks."""
    known = known.clone()
    plan = []
    while not known.all():
        keys = torch.where(known)[0]
        round_, pending = [], None
        for left, right in zip(keys[:-1].tolist(), keys[1:].tolist()):
            if pending is not None:
                if len(pending) + right - left <= max_tokens:
                    pending = torch.cat([pending, torch.arange(left + 1, right + 1)])
                    continue
                round_.append(pending)
                pending = None
            gap = right - left
            if gap == 1:
                continue
            if gap >= max_tokens - 1:
                round_.append(torch.linspace(left, right, max_tokens).round().long())
            else:
                pending = torch.arange(left, right + 1)
        if pending is not None:
            round_.append(pending)
        for frames in round_:
            known[frames] = True
        plan.append(round_)
    return plan


class SamplerOracle:
    def __init__(self, cfg: dict, model: Callable, randn: Callable = torch.randn,
                 randn_like: Callable = torch.randn_like):
        self.cfg = cfg
        d = cfg["diffusion"]
        self.diffusion = Diffusion(d, model, randn_like)
        self.randn, self.randn_like = randn, randn_like
        lat = cfg["latent"]
        self.x_shape = list(cfg["x_shape"])
        if lat["enabled"]:
            self.x_shape = list(lat["shape"]) if lat.get("shape") is not None else \
                [lat["num_channels"]] + [s // lat["downsampling_factor"][1] for s in self.x_shape[1:]]
        self.tds = lat["downsampling_factor"][0]
        self.timesteps = d["timesteps"]
        self.sampling_timesteps = d["sampling_timesteps"]
        self.clip_noise = d["clip_noise"]
        self.use_causal_mask = d["use_causal_mask"]
        self.max_tokens = (cfg["max_frames"] - 1) // self.tds + 1
        self.is_full_sequence = (cfg["noise_level"] == "random_uniform" and not cfg["fixed_context"]["enabled"]
                                 and not cfg["variable_context"]["enabled"])
        self.external_cond_dim = cfg["external_cond_dim"] * (cfg["frame_skip"] if cfg["external_cond_stack"] else 1)
        self.trace = None  # optional list; one dict of per-step tensors is appended per sampling step

    # base_pytorch_video_algo.py:635-664
    def process_conditions(self, cond, noise_levels=None):
        if cond is not None and "camera_pose_conditioning" in self.cfg:   # dfot_video_pose.py:64-110
            from .pose import ray_encoding
            cp = self.cfg["camera_pose_conditioning"]
            interp = None                                         # dfot_video_pose.py:73-81
            if self.cfg["tasks"]["prediction"]["history_guidance"]["name"] == "temporal":
                interp = noise_levels == self.timesteps - 1
            return ray_encoding(cond, self.x_shape[1], cp["normalize_by"], cp["bound"], cp["type"], interp_mask=interp)
        if cond is None or self.cfg["external_cond_processing"] is None:
            return cond
        assert self.cfg["external_cond_processing"] == "mask_first"
        keep = torch.ones_like(cond)
        keep[:, :1, : self.external_cond_dim] = 0
        return cond * keep

    def pad_to_max_tokens(self, y):
        if y is None or y.shape[1] >= self.max_tokens:
            return y
        tail = y[:, -1:].expand(-1, self.max_tokens - y.shape[1], *y.shape[2:])
        return torch.cat([y, tail], dim=1)

    def _bcast(self, a):
        return a.reshape(*a.shape, *([1] * len(self.x_shape)))

    # dfot_video.py:516-763
    def sample_sequence(self, batch_size: int, length: Optional[int], context: torch.Tensor,
                        context_mask: torch.Tensor, conditions=None, scheme: Optional[hg.Scheme] = None):
        if length is None:
            length = context.shape[1]
        if length > self.max_tokens:
            raise ValueError("length > max_tokens")
        horizon = length if self.use_causal_mask else self.max_tokens
        padding = horizon - length
        x = torch.clamp(self.randn((batch_size, horizon, *self.x_shape)), -self.clip_noise, self.clip_noise)
        context_mask = context_mask.long()
        if padding > 0:
            context = torch.cat([context, torch.zeros((batch_size, padding, *self.x_shape))], 1)
            context_mask = torch.cat([context_mask, -torch.ones((batch_size, padding), dtype=torch.long)], 1)
        if scheme is None:
            scheme = hg.scheme_from_config({"name": "conditional"}, self.timesteps)
        x = torch.where(self._bcast(context_mask) >= 1, context, x)

        S = scheduling_matrix(self.cfg["scheduling_matrix"], horizon - padding, padding, self.timesteps,
                              self.sampling_timesteps)
        S = S[:, None, :].repeat(1, batch_size, 1)
        if not self.is_full_sequence:
            S = torch.where(context_mask[None] >= 1, -1, S)
        same = (S[1:] - S[:-1] == 0).flatten(1).all(dim=1)
        S = S[int(torch.argmax((~same).float())):]

        for m in range(S.shape[0] - 1):
            frm, to = S[m], S[m + 1]
            context_mask = torch.where((context_mask == 0) & (frm == -1), 2, context_mask)
            x_prev = x.clone()
            if scheme.is_simple:
                xr, f, t, cond_mask = hg.simple_prepare(scheme, context_mask, x, frm, to, self.diffusion.q_sample)
                nfe = 1 if scheme.hist_weights[0] == 1 else 2
            else:
                assert (context_mask == context_mask[0]).all()
                tab = hg.branch_table(scheme, context_mask[0])
                xr, f, t, cond_mask, excluded = hg.full_prepare(
                    scheme, tab, context_mask, x, frm, to, self.diffusion.q_sample, self.is_full_sequence,
                    self.randn_like)
                nfe = tab.nfe
            cond = None
            if conditions is not None:
                cond = self.process_conditions(conditions.repeat_interleave(nfe, dim=0).clone(), f)
            x_new, model_out = self.diffusion.sample_step(xr, f, t, cond, cond_mask, return_model_out=True)
            composed = hg.simple_compose(scheme, x_new) if scheme.is_simple else hg.full_compose(tab, excluded, x_new)
            x = torch.where(self._bcast(context_mask) == 0, composed, x_prev)
            if self.trace is not None:
                self.trace.append(dict(model_in=xr, levels_from=f, levels_to=t, cond_mask=cond_mask,
                                       model_out=model_out, step_out=x_new, x_after=x.clone(),
                                       context_mask=context_mask.clone()))
        return x[:, :length] if padding > 0 else x

    # dfot_video.py:765-1008 (fork-only): the same window sampler over the refinement walk — a row whose LAST column's
    # level decreases is a denoising step, any other row re-noises every token from its level up to the next one
    def sample_sequence_refine(self, batch_size: int, length: Optional[int], context: torch.Tensor,
                               context_mask: torch.Tensor, conditions=None, scheme: Optional[hg.Scheme] = None):
        rs = self.cfg["refinement_sampling"]
        if length is None:
            length = context.shape[1]
        if length > self.max_tokens:
            raise ValueError("length > max_tokens")
        horizon = length if self.use_causal_mask else self.max_tokens
        padding = horizon - length
        x = torch.clamp(self.randn((batch_size, horizon, *self.x_shape)), -self.clip_noise, self.clip_noise)
        context_mask = context_mask.long()
        if padding > 0:
            context = torch.cat([context, torch.zeros((batch_size, padding, *self.x_shape))], 1)
            context_mask = torch.cat([context_mask, -torch.ones((batch_size, padding), dtype=torch.long)], 1)
        if scheme is None:
            scheme = hg.scheme_from_config({"name": "conditional"}, self.timesteps)
        x = torch.where(self._bcast(context_mask) >= 1, context, x)
        assert self.cfg["scheduling_matrix"] == "full_sequence", "Refining only support full_sequence scheduling matrix"
        S = refine_scheduling_matrix(horizon - padding, rs["goback_length"], rs["n_goback"], padding, self.timesteps,
                                     self.sampling_timesteps)
        S = S[:, None, :].repeat(1, batch_size, 1)
        S = torch.where(context_mask[None] >= 1, -1, S)                     # (:889-891) unconditionally here
        for m in range(S.shape[0] - 1):
            frm, to = S[m], S[m + 1]
            if frm[0, -1].item() > to[0, -1].item():
                context_mask = torch.where((context_mask == 0) & (frm == -1), 2, context_mask)
                x_prev = x.clone()
                assert scheme.is_simple and scheme.hist_weights[0] == 1, \
                    "the reference's refinement loop only runs with one branch (q_sample(context, to) at :983)"
                xr, f, t, cond_mask = hg.simple_prepare(scheme, context_mask, x, frm, to, self.diffusion.q_sample)
                cond = None
                if conditions is not None:
                    cond = self.process_conditions(conditions.clone(), f)
                x_new, model_out = self.diffusion.sample_step(xr, f, t, cond, cond_mask, return_model_out=True)
                composed = hg.simple_compose(scheme, x_new)
                xc_t = self.diffusion.q_sample(context, t)                   # (:983) draws noise; overwritten below
                composed = torch.where(self._bcast(context_mask) == 0, composed, xc_t)
                x = torch.where(self._bcast(context_mask) == 0, composed, x_prev)
                if self.trace is not None:
                    self.trace.append(dict(model_in=xr, levels_from=f, levels_to=t, cond_mask=cond_mask,
                                           model_out=model_out, step_out=x_new, x_after=x.clone(),
                                           context_mask=context_mask.clone()))
            else:
                x = self.diffusion.q_sample_from_x_k(x, frm, to)
        return x[:, :length] if padding > 0 else x

    def _window(self, *a):
        fn = self.sample_sequence_refine if self.cfg.get("refinement_sampling", {}).get("enabled") else self.sample_sequence
        return fn(*a)

    # dfot_video.py:362-514
    def predict_sequence(self, context, length=None, conditions=None, scheme=None, sliding_context_len=None):
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
            raise ValueError("sliding_context_len is expected to be >= length of initial context")
        chunk = self.cfg["chunk_size"] if self.use_causal_mask else self.max_tokens
        xs, cur = context, gt_len
        while cur < length:
            c = min(sliding_context_len, cur)
            h = min(length - cur, self.max_tokens - c)
            h = min(h, chunk) if chunk > 0 else h
            ctx = torch.cat([xs[:, -c:], torch.zeros((B, h, *self.x_shape))], 1)
            generated = cur - max(cur - c, gt_len)
            mask = torch.ones((B, c), dtype=torch.long)
            if generated > 0:
                mask[:, -generated:] = 2
            mask = torch.cat([mask, torch.zeros((B, h), dtype=torch.long)], 1)
            cond_len = c + h if self.use_causal_mask else self.max_tokens
            cond = None
            if conditions is not None:
                cond = conditions if self.cfg["external_cond_type"] == "label" else \
                    conditions[:, cur - c: cur - c + cond_len]
            new = self._window(B, c + h, ctx, mask, cond, scheme)
            xs = torch.cat([xs, new[:, -h:]], 1)
            cur = xs.shape[1]
        return xs

    # dfot_video.py:114-179
    def predict_videos(self, xs, n_context_tokens: int, conditions=None):
        task = self.cfg["tasks"]["prediction"]
        scheme = hg.scheme_from_config(task["history_guidance"], self.timesteps)
        out = xs.clone()
        density = task.get("keyframe_density") or 1
        if density > 1:
            raise ValueError("tasks.prediction.keyframe_density must be <= 1")
        T = out.shape[1]
        keys = torch.linspace(0, T - 1, round(density * T)).round().long()
        keys = torch.cat([torch.arange(n_context_tokens), keys]).unique()
        key_cond = None
        if conditions is not None:
            key_cond = conditions if self.cfg["external_cond_type"] == "label" else conditions[:, keys]
        pred = self.predict_sequence(out[:, :n_context_tokens], len(keys), key_cond, scheme,
                                     task.get("sliding_context_len") or self.max_tokens // 2)
        out[:, keys] = pred.to(out.dtype)
        if len(keys) < T:
            known = torch.zeros(out.shape[:2], dtype=torch.bool)
            known[:, keys] = True
            out = self.interpolate_videos(out, known, conditions)
        return out

    # dfot_video.py:181-360 (all chunks are processed, i.e. without quirk Q10)
    def interpolate_videos(self, context, context_mask=None, conditions=None):
        if context_mask is None:
            context_mask = torch.zeros(context.shape[:2], dtype=torch.bool)
            context_mask[:, [0, -1]] = True
        assert context_mask[:, [0, -1]].all()
        task = self.cfg["tasks"]["interpolation"]
        scheme = hg.scheme_from_config(task["history_guidance"], self.timesteps)
        plan = interpolation_plan(context_mask[0], self.max_tokens)
        xs, known = context.clone(), context_mask.clone()
        for round_ in plan:
            ctx = torch.cat([self.pad_to_max_tokens(xs[:, f]) for f in round_], 0)
            msk = torch.cat([self.pad_to_max_tokens(known[:, f]) for f in round_], 0)
            cnd = None
            if conditions is not None:
                cnd = torch.cat([conditions if self.cfg["external_cond_type"] == "label"
                                 else self.pad_to_max_tokens(conditions[:, f]) for f in round_], 0)
            mb = task.get("max_batch_size") or ctx.shape[0]
            outs = []
            for s in range(0, ctx.shape[0], mb):
                outs.append(self._window(min(mb, ctx.shape[0] - s), None, ctx[s:s + mb],
                                                 msk[s:s + mb].long(), None if cnd is None else cnd[s:s + mb], scheme))
            outs = torch.cat(outs, 0)
            for frames, pred in zip(round_, outs.chunk(len(round_), 0)):
                xs[:, frames] = pred[:, : len(frames)]
                known[:, frames] = True
        return xs
