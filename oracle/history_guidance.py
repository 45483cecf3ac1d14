"""ORACLE (test infrastructure): History Guidance — schemes, branch tables,
prepare and compose.

Restates algorithms/dfot/history_guidance.py:
  HistorySegment.to_noise_levels      :71-149
  HistoryGuidance schemes             :689-900, dispatch :635-653
  HistoryGuidanceManager.__enter__    :357-437   (branch table)
  HistoryGuidanceManager.prepare      :446-543
  HistoryGuidanceManager.compose      :545-568
  SimpleHistoryGuidanceManager        :903-982
Written table-first: a scheme is plain data, a step is (table, prepare, compose).
"""
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import torch

ALL = "all"


@dataclass
class Segment:
    time_indices: object = ALL                       # list[int] | "all"
    freq_ranges: list = field(default_factory=lambda: [ALL])
    freq_ranges_if_generated: Optional[list] = None

    def __post_init__(self):
        if self.freq_ranges is None:
            self.freq_ranges = [ALL]
        if self.freq_ranges_if_generated is None:
            self.freq_ranges_if_generated = self.freq_ranges


@dataclass
class Scheme:
    hist_segments: List[Segment]
    hist_weights: List[float]
    gen_segments: list = field(default_factory=lambda: [ALL])
    timesteps: int = 1000
    use_external_cond_guidance: bool = False

    @property
    def is_simple(self) -> bool:
        # history_guidance.py:646-653
        s = self.hist_segments[0]
        return (len(self.hist_weights) == 1 and len(s.freq_ranges) == 1 and s.freq_ranges[0] == ALL
                and s.freq_ranges_if_generated[0] == ALL)


def scheme_from_config(config: dict, timesteps: int = 1000) -> Scheme:
    """history_guidance.py:673-900; ``visualize`` is accepted and ignored."""
    cfg = dict(config)
    name = cfg.pop("name")
    cfg.pop("visualize", None)
    t = timesteps
    if name == "conditional":
        return Scheme([Segment()], [1], timesteps=t, use_external_cond_guidance=False)
    if name == "stabilized_conditional":
        lvl = cfg["stabilization_level"]
        return Scheme([Segment(ALL, [ALL], [(lvl, 1.0)])], [1], timesteps=t, use_external_cond_guidance=False)
    ext = cfg.pop("use_external_cond_guidance", True)
    if name == "vanilla":
        return Scheme([Segment()], [cfg["guidance_scale"]], timesteps=t, use_external_cond_guidance=ext)
    if name == "stabilized_vanilla":
        return Scheme([Segment(ALL, [ALL], [(cfg["stabilization_level"], 1.0)])], [cfg["guidance_scale"]],
                      timesteps=t, use_external_cond_guidance=ext)
    if name == "fractional":
        return Scheme([Segment(), Segment(ALL, [(cfg["freq_scale"], 1.0)])], [1, cfg["guidance_scale"] - 1],
                      timesteps=t, use_external_cond_guidance=ext)
    if name == "stabilized_fractional":
        return Scheme([Segment(ALL, [ALL], [(cfg["stabilization_level"], 1.0)]),
                       Segment(ALL, [(cfg["freq_scale"], 1.0)])], [1, cfg["guidance_scale"] - 1],
                      timesteps=t, use_external_cond_guidance=ext)
    if name == "temporal":
        return Scheme([Segment(time_indices=s) for s in cfg["hist_subsequences"]], cfg["hist_weights"],
                      gen_segments=cfg.get("gen_segments") or [ALL], timesteps=t, use_external_cond_guidance=ext)
    if name == "custom":
        def tup(fr):
            return None if fr is None else [tuple(x) if x != ALL else ALL for x in fr]
        segs = [Segment(s["time_indices"], tup(s["freq_ranges"]), tup(s.get("freq_ranges_if_generated")))
                for s in cfg["hist_segments"]]
        return Scheme(segs, cfg["hist_weights"], gen_segments=cfg.get("gen_segments") or [ALL], timesteps=t,
                      use_external_cond_guidance=ext)
    raise ValueError(name)


def _expand_ranges(freq_ranges: list, n: int) -> List[Tuple[float, float]]:
    # history_guidance.py:71-103
    fr = [(0.0, 1.0) if r == ALL else tuple(r) for r in freq_ranges]
    if len(fr) == n:
        return fr
    if len(fr) == 2:
        if n == 1:
            return [fr[1]]
        (a0, a1), (b0, b1) = fr
        return [(a0 + (b0 - a0) * i / (n - 1), a1 + (b1 - a1) * i / (n - 1)) for i in range(n)]
    if len(fr) == 1:
        return fr * n
    raise ValueError(f"history length {n} vs {len(fr)} freq_ranges")


def segment_levels(seg: Segment, generated: List[bool]) -> Tuple[tuple, tuple]:
    """history_guidance.py:105-149: (start_levels, end_levels) in [0,1] per history token."""
    n = len(generated)
    idx = list(range(n)) if seg.time_indices == ALL else list(seg.time_indices)
    assert all(t < n for t in idx)
    idx = [t if t >= 0 else n + t for t in idx]
    fr = _expand_ranges(seg.freq_ranges, len(idx))
    fr_gen = _expand_ranges(seg.freq_ranges_if_generated, len(idx))
    final = [(1.0, 1.0)] * n
    for i, t in enumerate(idx):
        final[t] = fr_gen[i] if generated[t] else fr[i]
    return tuple(zip(*final)) if n > 0 else ((), ())


@dataclass
class BranchTable:
    hist_indices: torch.Tensor      # int64 [hist_len]
    gen_indices: torch.Tensor       # int64 [gen_len]
    gen_mask: torch.Tensor          # bool  [g, T]
    hist_noise_levels: torch.Tensor  # int64 [h, hist_len]
    cond_mask: torch.Tensor         # bool  [h]
    weights: torch.Tensor           # fp32  [h]

    @property
    def nfe(self) -> int:
        return self.gen_mask.shape[0] * self.weights.shape[0]


def branch_table(scheme: Scheme, mask_row: torch.Tensor) -> BranchTable:
    """history_guidance.py:357-437 for one (batch-uniform) mask row."""
    hist_idx = torch.where(mask_row >= 1)[0]
    gen_idx = torch.where(mask_row == 0)[0]
    T, hl, gl = len(mask_row), len(hist_idx), len(gen_idx)
    segs = [list(range(gl)) if g == ALL else g for g in scheme.gen_segments]
    gen_mask = torch.zeros((len(segs), T), dtype=torch.bool)
    for i, g in enumerate(segs):
        gen_mask[i, gen_idx[g]] = True
    ext = scheme.use_external_cond_guidance
    table = OrderedDict()
    table[(1.0,) * hl + (ext,)] = 1.0
    generated = (mask_row[hist_idx] == 2).tolist()
    for seg, w in zip(scheme.hist_segments, scheme.hist_weights):
        start, end = segment_levels(seg, generated)
        ks, ke = start + (False,), end + (ext,)
        table[ks] = table.get(ks, 0.0) + w
        table[ke] = table.get(ke, 0.0) - w
    levels, flags, weights = [], [], []
    for key, w in table.items():
        if w == 0:
            continue
        levels.append(key[:-1])
        flags.append(key[-1])
        weights.append(w)
    # float32 arithmetic then truncation, :428-432
    lv = (torch.tensor(levels, dtype=torch.float32).reshape(len(levels), hl) * scheme.timesteps - 1).long()
    return BranchTable(hist_idx, gen_idx, gen_mask, lv, torch.tensor(flags, dtype=torch.bool),
                       torch.tensor(weights).float())


def _ext(a: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    return a.reshape(*a.shape, *([1] * (x.ndim - a.ndim)))


# --------------------------------------------------------------------------
# full manager, history_guidance.py:446-568
# --------------------------------------------------------------------------
def full_prepare(scheme: Scheme, tab: BranchTable, mask: torch.Tensor, x, frm, to,
                 replacement_fn: Callable, replacement_only: bool, randn_like: Callable):
    """rows ordered (b h g). Returns x, from, to, cond_mask, excluded_mask."""
    b, h, g = x.shape[0], tab.weights.shape[0], tab.gen_mask.shape[0]
    rep_h = lambda y: y.unsqueeze(1).repeat(1, h, *([1] * (y.ndim - 1))).clone()
    x, frm, to, m = rep_h(x), rep_h(frm), rep_h(to), rep_h(mask)
    if not replacement_only:
        frm[:, :, tab.hist_indices] = tab.hist_noise_levels
        to[:, :, tab.hist_indices] = tab.hist_noise_levels
    replace = (frm >= 0) & (m >= 1)
    noised = replacement_fn(x.flatten(0, 1), frm.flatten(0, 1)).reshape(x.shape)
    x = torch.where(_ext(replace, x), noised, x)
    rep_g = lambda y: y.flatten(0, 1).unsqueeze(1).repeat(1, g, *([1] * (y.ndim - 2))).clone()
    x, frm, to, m = rep_g(x), rep_g(frm), rep_g(to), rep_g(m)
    excluded = (~tab.gen_mask) & (m == 0)
    frm = torch.where(excluded, scheme.timesteps - 1, frm)
    to = torch.where(excluded, scheme.timesteps - 1, to)
    x = torch.where(_ext(excluded, x), randn_like(x), x)  # drawn always, :527-531
    flat = lambda y: y.flatten(0, 1)
    cond_mask = tab.cond_mask[None, :, None].expand(b, h, g).reshape(-1).clone()
    return flat(x), flat(frm), flat(to), cond_mask, excluded


def full_compose(tab: BranchTable, excluded: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    h, g = tab.weights.shape[0], tab.gen_mask.shape[0]
    x = x.reshape(-1, g, *x.shape[1:])                      # (b h) g t ...
    x = torch.where(_ext(excluded, x), torch.zeros_like(x), x)
    x = x.reshape(-1, h, *x.shape[1:])                      # b h g t ...
    x = torch.einsum("bhg...,h->bg...", x, tab.weights)
    x = x.sum(dim=1)
    denom = tab.gen_mask.long().sum(0).clamp(min=1)
    return x / denom.reshape(-1, *([1] * (x.ndim - 2)))


# --------------------------------------------------------------------------
# simple manager, history_guidance.py:929-982
# --------------------------------------------------------------------------
def simple_prepare(scheme: Scheme, mask: torch.Tensor, x, frm, to, replacement_fn: Callable):
    s = scheme.hist_weights[0]
    if s == 1:
        return x, frm, to, None
    rep = lambda y: y.unsqueeze(1).repeat(1, 2, *([1] * (y.ndim - 1))).clone()
    x, frm, to = rep(x), rep(frm), rep(to)
    hist = mask >= 1
    frm[:, 0] = torch.where(hist, scheme.timesteps - 1, frm[:, 0])
    to[:, 0] = torch.where(hist, scheme.timesteps - 1, to[:, 0])
    x[:, 0] = torch.where(_ext(hist, x[:, 0]), replacement_fn(x[:, 0], frm[:, 0]), x[:, 0])
    b = x.shape[0]
    cond_mask = torch.tensor([True, False]).repeat(b) if scheme.use_external_cond_guidance else None
    return x.flatten(0, 1), frm.flatten(0, 1), to.flatten(0, 1), cond_mask


def simple_compose(scheme: Scheme, x: torch.Tensor) -> torch.Tensor:
    s = scheme.hist_weights[0]
    if s == 1:
        return x
    x = x.reshape(-1, 2, *x.shape[1:])
    return x[:, 1] * s - x[:, 0] * (s - 1)
