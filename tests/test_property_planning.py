"""Property tests (hypothesis, CPU): the product's integer planning must equal the oracle's — which is pinned against the
reference-generated goldens — on configurations far outside the fixed fixtures: scheduling matrices for random
(kind, horizon, padding, steps), noise-level tables, history-guidance branch tables for random masks and schemes, and
interpolation plans for random keyframe sets.  Integers are compared bit for bit."""
import numpy as np
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from dfot_b200.algorithms.dfot import DFoTVideo
from dfot_b200.algorithms.dfot.dfot_video import interpolation_plan as product_plan
from dfot_b200.algorithms.dfot.history_guidance import HistoryGuidance
from oracle import history_guidance as ohg
from oracle import schedule
from oracle.cases import algorithm_cfg
from oracle.sampler import interpolation_plan as oracle_plan

SETTINGS = dict(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
_ALGOS = {}


def _algo(kind, steps, max_frames):
    key = (kind, steps, max_frames)
    if key not in _ALGOS:
        _ALGOS[key] = DFoTVideo(algorithm_cfg(**{
            "backbone.hidden_size": 64, "backbone.depth": 1, "backbone.num_heads": 1, "x_shape": [4, 8, 8],
            "scheduling_matrix": kind, "diffusion.sampling_timesteps": steps, "max_frames": max_frames}))
    return _ALGOS[key]


@settings(**SETTINGS)
@given(kind=st.sampled_from(["full_sequence", "autoregressive", "interleaved"]), horizon=st.integers(1, 12),
       padding=st.integers(0, 3), steps=st.sampled_from([1, 2, 3, 4, 7, 10, 25, 50]))
def test_scheduling_matrix_equals_oracle(kind, horizon, padding, steps):
    algo = _algo(kind, steps, horizon + padding)
    got = algo._generate_scheduling_matrix(horizon, padding)
    want = schedule.scheduling_matrix(kind, horizon, padding, 1000, steps)
    assert got.dtype == torch.int64 and got.tolist() == want.tolist()
    lv = algo.diffusion_model.ddim_idx_to_noise_level(torch.arange(steps + 1))
    assert lv.tolist() == schedule.ddim_idx_to_noise_level(torch.arange(steps + 1), 1000, steps).tolist()


def _masks():
    # [ground-truth context][generated context][to generate][padding], at least one frame to generate
    return st.tuples(st.integers(0, 4), st.integers(0, 4), st.integers(1, 5), st.integers(0, 3)).filter(
        lambda t: t[0] + t[1] >= 1).map(lambda t: [1] * t[0] + [2] * t[1] + [0] * t[2] + [-1] * t[3])


_levels = st.sampled_from([0.0, 0.001, 0.02, 0.05, 0.1, 0.3, 0.5, 1.0])
_schemes = st.one_of(
    st.just(dict(name="conditional")),
    st.builds(lambda s: dict(name="vanilla", guidance_scale=s), st.sampled_from([1.0, 1.5, 2.0, 4.0])),
    st.builds(lambda l: dict(name="stabilized_conditional", stabilization_level=l), _levels),
    st.builds(lambda s, l: dict(name="stabilized_vanilla", guidance_scale=s, stabilization_level=l),
              st.sampled_from([1.5, 4.0]), _levels),
    st.builds(lambda s, f: dict(name="fractional", guidance_scale=s, freq_scale=f), st.sampled_from([2.0, 4.0]),
              st.sampled_from([0.1, 0.3, 0.7])),
    st.builds(lambda s, f, l: dict(name="stabilized_fractional", guidance_scale=s, freq_scale=f, stabilization_level=l),
              st.sampled_from([2.0, 4.0]), st.sampled_from([0.1, 0.3]), _levels))


@settings(**SETTINGS)
@given(mask=_masks(), scheme=_schemes)
def test_history_guidance_tables_equal_oracle(mask, scheme):
    hgd = HistoryGuidance.from_config(dict(scheme, visualize=False), timesteps=1000)
    osch = ohg.scheme_from_config(scheme, 1000)
    assert hgd.is_simple == osch.is_simple
    if hgd.is_simple:
        return
    got, want = hgd.branch_table(np.array(mask)), ohg.branch_table(osch, torch.tensor(mask))
    assert got.num_hist * got.num_gen == want.nfe
    assert got.hist_indices.tolist() == want.hist_indices.tolist()
    assert got.gen_indices.tolist() == want.gen_indices.tolist()
    assert got.gen_mask.astype(int).tolist() == want.gen_mask.long().tolist()
    assert got.hist_noise_levels.tolist() == want.hist_noise_levels.tolist()
    assert got.cond_mask.astype(int).tolist() == want.cond_mask.long().tolist()
    assert got.weights.tolist() == want.weights.tolist()


@settings(**SETTINGS)
@given(n_frames=st.integers(3, 120), max_tokens=st.integers(3, 16), data=st.data())
def test_interpolation_plan_equals_oracle(n_frames, max_tokens, data):
    inner = data.draw(st.lists(st.integers(1, n_frames - 2), unique=True, max_size=min(12, n_frames - 2))) if n_frames > 2 else []
    known = np.zeros(n_frames, dtype=bool)
    known[[0, n_frames - 1] + inner] = True
    got = product_plan(known.copy(), max_tokens)
    want = oracle_plan(torch.from_numpy(known.copy()), max_tokens)
    assert len(got) == len(want)
    for rg, rw in zip(got, want):
        assert [np.asarray(c).tolist() for c in rg] == [torch.as_tensor(c).tolist() for c in rw]
    filled = known.copy()
    for chunks in got:
        for c in chunks:
            assert len(c) <= max_tokens and filled[c[0]] and filled[c[-1]]      # every chunk is anchored at both ends
            filled[np.asarray(c)] = True
    assert filled.all()                                                            # the plan reaches every frame


@settings(**SETTINGS)
@given(horizon=st.integers(1, 8), padding=st.integers(0, 3), steps=st.sampled_from([2, 3, 5, 8, 10, 25, 50]),
       goback_length=st.integers(1, 20), n_goback=st.integers(0, 4))
def test_refine_scheduling_matrix_equals_oracle(horizon, padding, steps, goback_length, n_goback):
    """The refinement walk (base_pytorch_video_algo.py:949-976): product == oracle, every row transition moves the walk
    by exactly one DDIM index, it ends at level -1, and it has steps + 1 + 2 * goback_length * n_goback * (#go-back
    indices) rows."""
    algo = _algo("full_sequence", steps, horizon + padding)
    got = algo._generate_refine_scheduling_matrix(horizon, goback_length, n_goback, padding)
    want = schedule.refine_scheduling_matrix(horizon, goback_length, n_goback, padding, 1000, steps)
    assert got.dtype == torch.int64 and got.tolist() == want.tolist()
    n_idx = len(range(1, steps - goback_length, goback_length))
    assert got.shape == (steps + 1 + 2 * goback_length * n_goback * n_idx, horizon + padding)
    assert (got[-1, :horizon] == -1).all() and (got[:, horizon:] == 999).all()
    table = schedule.ddim_idx_to_noise_level(torch.arange(steps + 1), 1000, steps).tolist()
    idx = [table.index(v) for v in got[:, 0].tolist()]
    assert all(abs(a - b) == 1 for a, b in zip(idx[:-1], idx[1:])) and max(idx) <= steps


@settings(**SETTINGS)
@given(data=st.data(), rows=st.integers(1, 3), frames=st.integers(1, 6))
def test_renoise_records_equal_oracle_arithmetic(data, rows, frames):
    """`renoise_table` (the K4 records of q_sample_from_x_k, discrete_diffusion.py:252-260) for random level pairs —
    context (-1), pad (999) and ordinary levels, also downward pairs — against the oracle's fp32 arithmetic."""
    from oracle.cases import continuous_overrides
    from oracle.diffusion import Diffusion
    key = ("renoise", 0, 0)
    if key not in _ALGOS:
        cfg = algorithm_cfg(**{"backbone.hidden_size": 64, "backbone.depth": 1, "backbone.num_heads": 1, "x_shape": [4, 8, 8],
                               **continuous_overrides()})
        _ALGOS[key] = (DFoTVideo(cfg), Diffusion(cfg["diffusion"], None))
    algo, orc = _ALGOS[key]
    lv = st.sampled_from([-1, 0, 19, 165, 332, 499, 665, 832, 998, 999])
    cur = np.array(data.draw(st.lists(st.lists(lv, min_size=frames, max_size=frames), min_size=rows, max_size=rows)))
    nxt = np.array(data.draw(st.lists(st.lists(lv, min_size=frames, max_size=frames), min_size=rows, max_size=rows)))
    prep = algo.diffusion_model.renoise_table(cur, nxt)
    x, n = torch.ones((rows, frames, 1, 1, 1)), torch.full((rows, frames, 1, 1, 1), 2.0)
    ref = orc.q_sample_from_x_k(x, torch.from_numpy(cur), torch.from_numpy(nxt), n).reshape(rows, frames).numpy()
    got = prep["qa"] * 1.0 + prep["qb"] * 2.0
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.abs(got[ok] - ref[ok]).max(initial=0.0) <= 1e-6
    assert (prep["noise_row"] == np.arange(rows)[:, None]).all()
