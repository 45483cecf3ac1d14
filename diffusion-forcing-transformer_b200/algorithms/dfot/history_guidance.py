"""History Guidance for the B200 sampler — same public surface as the reference's
algorithms/dfot/history_guidance.py (``HistoryGuidance.from_config`` / scheme classmethods :673-900,
``hg(mask)`` context manager with ``.nfe`` / ``.prepare`` / ``.compose`` :311-568, :903-982).

Difference in *how*: a guidance scheme is compiled on the host into a branch table (numpy) — the
reference rebuilds it on the device, with ``.tolist()`` syncs, every sampling step — and ``prepare`` /
``compose`` are single launches of the fused K4 kernel instead of ~15 ATen passes.  The sampler in
dfot_video.py goes one step further and plans whole windows with ``plan_step`` (no per-step host work).
The matplotlib visualiser (:169-308) is debug tooling and out of scope: ``visualize`` is accepted and ignored.
"""
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from dfot_b200 import ops
from dfot_b200.config import to_container
from . import sampling_plan as sp

ALL = "all"


class HistorySegment:
    """A patch selection on the (time x frequency) grid of the history (history_guidance.py:21-166)."""

    def __init__(self, time_indices=ALL, freq_ranges=None, freq_ranges_if_generated=None):
        self.time_indices = time_indices
        self.freq_ranges = freq_ranges if freq_ranges is not None else [ALL]
        self.freq_ranges_if_generated = self.freq_ranges if freq_ranges_if_generated is None else freq_ranges_if_generated

    @staticmethod
    def _spread(ranges, n: int) -> List[Tuple[float, float]]:
        rs = [(0.0, 1.0) if r == ALL else (r[0], r[1]) for r in ranges]
        if len(rs) == n:
            return rs
        if len(rs) == 2:
            if n == 1:
                return [rs[1]]
            (s0, e0), (s1, e1) = rs
            return [(s0 + (s1 - s0) * i / (n - 1), e0 + (e1 - e0) * i / (n - 1)) for i in range(n)]
        if len(rs) == 1:
            return rs * n
        raise ValueError(f"The length of the history is {n}, but the length of freq_ranges is {len(rs)}.")

    def to_noise_levels(self, generated: Sequence[bool]) -> Tuple[tuple, tuple]:
        """(start, end) noise levels in [0, 1] per history token; tokens outside the segment are (1, 1)."""
        n = len(generated)
        chosen = list(range(n)) if self.time_indices == ALL else list(self.time_indices)
        assert all(t < n for t in chosen), "time_indices should be between 0 and hist_len."
        chosen = [t if t >= 0 else n + t for t in chosen]
        gt = self._spread(self.freq_ranges, len(chosen))
        gen = self._spread(self.freq_ranges_if_generated, len(chosen))
        lv = [(1.0, 1.0)] * n
        for i, t in enumerate(chosen):
            lv[t] = gen[i] if generated[t] else gt[i]
        if n == 0:
            return (), ()
        return tuple(s for s, _ in lv), tuple(e for _, e in lv)

    @classmethod
    def full(cls):
        return cls(time_indices=ALL, freq_ranges=[ALL])

    @classmethod
    def partial_constant(cls, start_freq: float, end_freq: float):
        return cls(time_indices=ALL, freq_ranges=[(start_freq, end_freq)])

    @classmethod
    def partial_linear(cls, first_range, last_range):
        return cls(time_indices=ALL, freq_ranges=[first_range, last_range])


@dataclass
class BranchTable:
    """Compiled guidance for one (batch-uniform) mask row."""
    hist_indices: np.ndarray        # int64 [hist_len]
    gen_indices: np.ndarray         # int64 [gen_len]
    gen_mask: np.ndarray            # bool [g, T]
    hist_noise_levels: np.ndarray   # int64 [h, hist_len]
    cond_mask: np.ndarray           # bool [h]
    weights: np.ndarray             # f32 [h]

    @property
    def num_hist(self) -> int:
        return int(self.weights.shape[0])

    @property
    def num_gen(self) -> int:
        return int(self.gen_mask.shape[0])


class HistoryGuidance:
    def __init__(self, hist_segments: List[HistorySegment], hist_weights: List[float], gen_segments=None,
                 timesteps: int = 1000, use_external_cond_guidance: bool = False, visualize: bool = True):
        assert len(hist_segments) == len(hist_weights), \
            f"Length of hist_segments and hist_weights should be the same, but got {len(hist_segments)} and {len(hist_weights)}."
        self.hist_segments = hist_segments
        self.hist_weights = hist_weights
        self.gen_segments = [ALL] if gen_segments is None else gen_segments
        assert len(self.gen_segments) > 0, "At least one gen_segment should be provided."
        self.timesteps = timesteps
        self.use_external_cond_guidance = use_external_cond_guidance

    # ---- dispatch (history_guidance.py:635-653)
    @property
    def is_simple(self) -> bool:
        s = self.hist_segments[0]
        return (len(self.hist_weights) == 1 and len(s.freq_ranges) == 1 and s.freq_ranges[0] == ALL
                and s.freq_ranges_if_generated[0] == ALL)

    def __call__(self, mask: torch.Tensor):
        return SimpleHistoryGuidanceManager(self, mask) if self.is_simple else HistoryGuidanceManager(self, mask)

    def log(self, logger=None):  # visualiser is out of scope
        return None

    # ---- host compilation (history_guidance.py:357-437)
    def branch_table(self, mask_row: np.ndarray) -> BranchTable:
        mask_row = np.asarray(mask_row)
        hist_idx = np.nonzero(mask_row >= 1)[0]
        gen_idx = np.nonzero(mask_row == 0)[0]
        T = mask_row.shape[0]
        segs = [list(range(len(gen_idx))) if g == ALL else g for g in self.gen_segments]
        gen_mask = np.zeros((len(segs), T), dtype=bool)
        for i, g in enumerate(segs):
            gen_mask[i, gen_idx[g]] = True
        ext = self.use_external_cond_guidance
        generated = [bool(v) for v in (mask_row[hist_idx] == 2)]
        acc: Dict[tuple, float] = {(1.0,) * len(hist_idx) + (ext,): 1.0}   # the unconditional score
        for seg, w in zip(self.hist_segments, self.hist_weights):
            start, end = seg.to_noise_levels(generated)
            for key, dw in ((start + (False,), w), (end + (ext,), -w)):
                acc[key] = acc.get(key, 0.0) + dw
        rows = [(k[:-1], k[-1], w) for k, w in acc.items() if w != 0]
        # float32 multiply-subtract then truncation, exactly as (tensor * timesteps - 1).long() at :428-432
        lv32 = np.array([r[0] for r in rows], dtype=np.float32).reshape(len(rows), len(hist_idx))
        levels = (lv32 * np.float32(self.timesteps) - np.float32(1)).astype(np.int64)
        return BranchTable(hist_idx, gen_idx, gen_mask, levels, np.array([r[1] for r in rows], dtype=bool),
                           np.array([r[2] for r in rows], dtype=np.float32))

    def plan_step(self, tb: sp.HostTables, mask: np.ndarray, frm: np.ndarray, to: np.ndarray, replacement_only: bool,
                  continuous: bool, precond_scale: float) -> sp.StepPlan:
        """Compile one sampling step for the fused kernel.  mask/frm/to: int64 [B, T] (host)."""
        B, T = mask.shape
        tmax = self.timesteps - 1
        if self.is_simple:
            s = float(self.hist_weights[0])
            if s == 1:
                nfe, weights = 1, np.ones((1,), np.float32)
                f, t = frm[:, None].copy(), to[:, None].copy()
                mode = np.zeros((B, 1, T), np.int32)
                cond_mask, n_hist = None, 0
            else:
                nfe, weights = 2, np.array([-(s - 1), s], np.float32)
                f, t = np.repeat(frm[:, None], 2, 1), np.repeat(to[:, None], 2, 1)
                hist = mask >= 1
                f[:, 0][hist] = tmax
                t[:, 0][hist] = tmax
                mode = np.zeros((B, 2, T), np.int32)
                mode[:, 0][hist] = sp.MODE_QSAMPLE
                cond_mask = np.tile(np.array([True, False]), B) if self.use_external_cond_guidance else None
                n_hist = B
            noise_row = np.broadcast_to(np.arange(B, dtype=np.int32)[:, None, None], (B, nfe, T))
            w = np.broadcast_to(weights[None, :, None], (B, nfe, T)).astype(np.float32)
            draws_excl = False
        else:
            assert (mask == mask[0]).all(), "`mask` should be the same across the batch to use history guidance."
            tab = self.branch_table(mask[0])
            h, g = tab.num_hist, tab.num_gen
            nfe = h * g
            f = np.repeat(frm[:, None], h, 1)           # [B, h, T]
            t = np.repeat(to[:, None], h, 1)
            if not replacement_only:
                f[:, :, tab.hist_indices] = tab.hist_noise_levels[None]
                t[:, :, tab.hist_indices] = tab.hist_noise_levels[None]
            replace = (f >= 0) & (mask[:, None, :] >= 1)   # [B, h, T]
            expand = lambda y: np.repeat(y[:, :, None], g, 2)  # → [B, h, g, T]
            f, t, replace = expand(f), expand(t), expand(replace)
            excluded = (~tab.gen_mask)[None, None] & (mask[:, None, None, :] == 0)
            f = np.where(excluded, tmax, f)
            t = np.where(excluded, tmax, t)
            mode = np.where(excluded, sp.MODE_NOISE, np.where(replace, sp.MODE_QSAMPLE, sp.MODE_COPY)).astype(np.int32)
            noise_row = np.broadcast_to((np.arange(B)[:, None] * h + np.arange(h)[None, :]).astype(np.int32)[:, :, None, None],
                                        (B, h, g, T))
            denom = np.clip(tab.gen_mask.sum(0), 1, None).astype(np.float32)      # [T]
            w = np.where(excluded, 0.0, tab.weights[None, :, None, None] / denom[None, None, None, :]).astype(np.float32)
            f, t, mode, noise_row, w = (y.reshape(B, nfe, T) for y in (f, t, mode, noise_row, w))
            cond_mask = np.broadcast_to(tab.cond_mask[None, :, None], (B, h, g)).reshape(-1).copy()
            n_hist, draws_excl = B * h, True
        f2, t2 = f.reshape(B * nfe, T), t.reshape(B * nfe, T)
        generate = np.repeat((mask == 0).astype(np.int32)[:, None], nfe, 1).reshape(B * nfe, T)
        update = sp.step_update_table(tb, f2, t2, w.reshape(B * nfe, T), generate)
        prep = np.zeros((B * nfe, T), dtype=sp.PREPARE_DTYPE)
        prep["mode"] = mode.reshape(B * nfe, T)
        prep["noise_row"] = noise_row.reshape(B * nfe, T)
        kq = f2  # q_sample level = the branch's from-level (k = -1 indexes the last entry; masked out by mode)
        prep["qa"] = tb.sqrt_alphas_cumprod[kq]
        prep["qb"] = tb.sqrt_one_minus_alphas_cumprod[kq]
        return sp.StepPlan(nfe=nfe, levels=sp.model_levels(tb, f2, continuous, precond_scale), cond_mask=cond_mask,
                           update=update, prepare=prep, n_hist_rows=n_hist, draws_excluded_noise=draws_excl,
                           context_mask=mask.copy(), levels_from=f2, levels_to=t2)

    # ---- configuration surface (history_guidance.py:673-900)
    @classmethod
    def from_config(cls, config, timesteps: int = 1000) -> "HistoryGuidance":
        config = dict(to_container(config))
        name = config.pop("name")
        return getattr(cls, name)(**config, timesteps=timesteps)

    @classmethod
    def conditional(cls, timesteps: int = 1000, visualize: bool = True):
        return cls([HistorySegment.full()], [1], timesteps=timesteps, use_external_cond_guidance=False)

    @classmethod
    def stabilized_conditional(cls, stabilization_level: float, timesteps: int = 1000, visualize: bool = True):
        seg = HistorySegment(ALL, [ALL], [(stabilization_level, 1.0)])
        return cls([seg], [1], timesteps=timesteps, use_external_cond_guidance=False)

    @classmethod
    def vanilla(cls, guidance_scale: float, timesteps: int = 1000, use_external_cond_guidance: bool = True,
                visualize: bool = True):
        return cls([HistorySegment.full()], [guidance_scale], timesteps=timesteps,
                   use_external_cond_guidance=use_external_cond_guidance)

    @classmethod
    def stabilized_vanilla(cls, guidance_scale: float, stabilization_level: float, timesteps: int = 1000,
                           use_external_cond_guidance: bool = True, visualize: bool = True):
        seg = HistorySegment(ALL, [ALL], [(stabilization_level, 1.0)])
        return cls([seg], [guidance_scale], timesteps=timesteps, use_external_cond_guidance=use_external_cond_guidance)

    @classmethod
    def fractional(cls, guidance_scale: float, freq_scale: float, timesteps: int = 1000,
                   use_external_cond_guidance: bool = True, visualize: bool = True):
        return cls([HistorySegment.full(), HistorySegment.partial_constant(freq_scale, 1.0)], [1, guidance_scale - 1],
                   timesteps=timesteps, use_external_cond_guidance=use_external_cond_guidance)

    @classmethod
    def stabilized_fractional(cls, guidance_scale: float, freq_scale: float, stabilization_level: float,
                              timesteps: int = 1000, use_external_cond_guidance: bool = True, visualize: bool = True):
        seg = HistorySegment(ALL, [ALL], [(stabilization_level, 1.0)])
        return cls([seg, HistorySegment.partial_constant(freq_scale, 1.0)], [1, guidance_scale - 1],
                   timesteps=timesteps, use_external_cond_guidance=use_external_cond_guidance)

    @classmethod
    def temporal(cls, hist_subsequences, hist_weights, gen_segments=None, timesteps: int = 1000,
                 use_external_cond_guidance: bool = True, visualize: bool = True):
        return cls([HistorySegment(time_indices=s) for s in hist_subsequences], hist_weights,
                   gen_segments=gen_segments if gen_segments is not None else [ALL], timesteps=timesteps,
                   use_external_cond_guidance=use_external_cond_guidance)

    @classmethod
    def custom(cls, hist_segments: List[Dict[str, Any]], hist_weights, gen_segments=None, timesteps: int = 1000,
               use_external_cond_guidance: bool = True, visualize: bool = True):
        def tup(fr):
            return None if fr is None else [tuple(x) if x != ALL else ALL for x in fr]
        segs = [HistorySegment(s["time_indices"], tup(s["freq_ranges"]), tup(s.get("freq_ranges_if_generated")))
                for s in hist_segments]
        return cls(segs, hist_weights, gen_segments=gen_segments if gen_segments is not None else [ALL],
                   timesteps=timesteps, use_external_cond_guidance=use_external_cond_guidance)


# --------------------------------------------------------------------------------------------------------
# Context managers with the reference's per-step API.  Each call is ONE launch of the fused kernel.
# They need a ``diffusion`` object for the tables; ``replacement_fn`` is accepted for signature
# compatibility and must be that object's bound ``q_sample``.
# --------------------------------------------------------------------------------------------------------
class _ManagerBase:
    def __init__(self, history_guidance: HistoryGuidance, mask: torch.Tensor):
        self.history_guidance = history_guidance
        self.mask = mask
        self.device = mask.device
        self._plan: Optional[sp.StepPlan] = None

    def __enter__(self):
        self._mask_host = self.mask.detach().cpu().numpy().astype(np.int64)   # one small D2H per step (API path only)
        self._enter()
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        return None

    def prepare(self, x: torch.Tensor, from_noise_levels: torch.Tensor, to_noise_levels: torch.Tensor,
                replacement_fn: Callable, replacement_only: bool = False):
        diffusion = getattr(replacement_fn, "__self__", None)
        if diffusion is None or not hasattr(diffusion, "host_tables"):
            raise RuntimeError("dfot_b200: replacement_fn must be the bound q_sample of a dfot_b200 diffusion model")
        hgd = self.history_guidance
        frm = from_noise_levels.detach().cpu().numpy().astype(np.int64)
        to = to_noise_levels.detach().cpu().numpy().astype(np.int64)
        plan = hgd.plan_step(diffusion.host_tables, self._mask_host, frm, to, replacement_only,
                             diffusion.is_continuous, diffusion.precond_scale)
        self._plan = plan
        B, T = frm.shape
        if plan.nfe == 1 and hgd.is_simple:
            return x, from_noise_levels, to_noise_levels, None
        nfe = plan.nfe
        x = x.contiguous().float()
        noise_hist = noise_excl = None
        if plan.n_hist_rows:
            noise_hist = torch.clamp(diffusion.randn((plan.n_hist_rows, *x.shape[1:]), x.device),
                                     -diffusion.clip_noise, diffusion.clip_noise)
        if plan.draws_excluded_noise:
            noise_excl = diffusion.randn((B * nfe, *x.shape[1:]), x.device)
        out = torch.empty((B * nfe, *x.shape[1:]), dtype=torch.float32, device=x.device)
        ops.sampler_step_hg(x.clone(), None, out, None, sp.to_device_bytes(plan.prepare, x.device), None, noise_hist,
                            noise_excl, B, nfe, T, max_noise_row=plan.n_hist_rows - 1 if plan.n_hist_rows else None)
        dev = x.device
        cond_mask = None if plan.cond_mask is None else torch.from_numpy(plan.cond_mask).to(dev)
        return (out, torch.from_numpy(plan.levels_from).to(dev), torch.from_numpy(plan.levels_to).to(dev), cond_mask)

    def compose(self, x: torch.Tensor) -> torch.Tensor:
        plan = self._plan
        if plan is None:
            raise RuntimeError("compose() called before prepare()")
        if plan.nfe == 1 and self.history_guidance.is_simple:
            return x
        nfe = plan.nfe
        B, T = x.shape[0] // nfe, x.shape[1]
        upd = plan.update.copy()
        upd["a"], upd["b"], upd["sigma"], upd["clip"], upd["generate"] = 0.0, 1.0, 0.0, 0.0, 1
        out = torch.zeros((B, *x.shape[1:]), dtype=torch.float32, device=x.device)
        ops.sampler_step_hg(out, x.contiguous().float(), None, sp.to_device_bytes(upd, x.device), None, None, None, None,
                            B, nfe, T)
        return out


class HistoryGuidanceManager(_ManagerBase):
    def _enter(self):
        m = self._mask_host
        assert (m == m[0]).all(), "`mask` should be the same across the batch to use history guidance."
        tab = self.history_guidance.branch_table(m[0])
        self.table = tab
        self.hist_indices = torch.from_numpy(tab.hist_indices).to(self.device)
        self.gen_indices = torch.from_numpy(tab.gen_indices).to(self.device)
        self.gen_mask = torch.from_numpy(tab.gen_mask).to(self.device)
        self.hist_noise_levels = torch.from_numpy(tab.hist_noise_levels).to(self.device)
        self.cond_mask = torch.from_numpy(tab.cond_mask).to(self.device)
        self.weights = torch.from_numpy(tab.weights).to(self.device)
        self.num_gen, self.num_hist = tab.num_gen, tab.num_hist

    @property
    def nfe(self) -> int:
        return self.num_gen * self.num_hist


class SimpleHistoryGuidanceManager(_ManagerBase):
    def _enter(self):
        self.guidance_scale = self.history_guidance.hist_weights[0]

    @property
    def nfe(self) -> int:
        return 1 if self.history_guidance.hist_weights[0] == 1 else 2
