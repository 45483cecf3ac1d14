"""Host-side logic of the product (no GPU): scheduling matrices, HG branch tables, interpolation plans,
diffusion tables, state-dict keys — bit-exact against the reference-generated goldens — and the window
planner driven end-to-end on CPU with the K4 *contract emulation* + the oracle backbone."""
import json
import os

import numpy as np
import pytest
import torch

from dfot_b200 import ops
from dfot_b200.algorithms.dfot import DFoTVideo
from dfot_b200.algorithms.dfot.dfot_video import interpolation_plan
from dfot_b200.algorithms.dfot.history_guidance import HistoryGuidance
from helpers import GOLDEN, MATRIX_COMBOS, NoiseBank, build_oracle, build_product, case_names, load_case, matrix_combo_model
from oracle.cases import algorithm_cfg, continuous_overrides
import k4_emulation

with open(os.path.join(GOLDEN, "integers.json")) as f:
    INTS = json.load(f)


def _tiny(**over):
    base = {"backbone.hidden_size": 64, "backbone.depth": 1, "backbone.num_heads": 1, "x_shape": [4, 8, 8]}
    base.update(over)
    return algorithm_cfg(**base)


def test_scheduling_matrices_bit_exact():
    for rec in INTS["scheduling_matrices"]:
        algo = DFoTVideo(_tiny(**{"scheduling_matrix": rec["kind"], "diffusion.sampling_timesteps": rec["steps"],
                                  "max_frames": rec["horizon"] + rec["padding"]}))
        m = algo._generate_scheduling_matrix(rec["horizon"], rec["padding"])
        assert m.dtype == torch.int64 and m.tolist() == rec["matrix"], rec["kind"]


def test_ddim_levels_bit_exact():
    for rec in INTS["ddim_levels"]:
        algo = DFoTVideo(_tiny(**{"diffusion.sampling_timesteps": rec["steps"]}))
        lv = algo.diffusion_model.ddim_idx_to_noise_level(torch.arange(rec["steps"] + 1))
        assert lv.tolist() == rec["levels"]


def test_hg_branch_tables_bit_exact():
    for rec in INTS["hg_branch_tables"]:
        hgd = HistoryGuidance.from_config(dict(rec["scheme"], visualize=False), timesteps=1000)
        mask = np.array(rec["mask"])
        if "error" in rec:
            with pytest.raises((AssertionError, IndexError)):
                hgd.branch_table(mask)
            continue
        assert hgd.is_simple == (rec["manager"] == "SimpleHistoryGuidanceManager")
        if hgd.is_simple:
            continue
        tab = hgd.branch_table(mask)
        assert tab.num_hist * tab.num_gen == rec["nfe"]
        assert tab.hist_indices.tolist() == rec["hist_indices"]
        assert tab.gen_indices.tolist() == rec["gen_indices"]
        assert tab.gen_mask.astype(int).tolist() == rec["gen_mask"]
        assert tab.hist_noise_levels.tolist() == rec["hist_noise_levels"], rec
        assert tab.cond_mask.astype(int).tolist() == rec["cond_mask"]
        assert tab.weights.tolist() == rec["weights"]


def test_interpolation_plans_bit_exact():
    for rec in INTS["interpolation_calls"]:
        known = np.zeros(rec["n_frames"], dtype=bool)
        known[rec["keyframes"]] = True
        plan = interpolation_plan(known, rec["max_tokens"])
        assert len(plan) == len(rec["calls"])
        k = known.copy()
        for chunks, call in zip(plan, rec["calls"]):
            rows = []
            for c in chunks:
                r = k[c].astype(int).tolist()
                rows.append(r + [r[-1]] * (rec["max_tokens"] - len(r)))
            assert rows == call["mask"]
            for c in chunks:
                k[c] = True


def test_diffusion_buffers_bit_exact():
    gold = np.load(os.path.join(GOLDEN, "schedules.npz"))
    for tag, over in [("cosine", {}), ("continuous", continuous_overrides()),
                      ("sigmoid_zt", {"diffusion.beta_schedule": "sigmoid", "diffusion.schedule_fn_kwargs": {}}),
                      ("cosine_shift", {"diffusion.schedule_fn_kwargs": dict(shift=0.5)})]:
        dm = DFoTVideo(_tiny(**over)).diffusion_model
        for name in ["alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "logsnr"]:
            if f"{tag}.{name}" in gold:
                assert np.array_equal(getattr(dm, name).numpy(), gold[f"{tag}.{name}"]), (tag, name)


@pytest.mark.parametrize("name", case_names())
def test_state_dict_keys_match_reference(name):
    meta, _, weights = load_case(name)
    algo = build_product(meta["cfg"])
    sd = {"diffusion_model.model." + k: v for k, v in weights.items()}
    sd["data_mean"], sd["data_std"] = algo.data_mean, algo.data_std
    res = algo.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_ops_fail_loudly_without_cuda():
    x = torch.zeros(2, 4, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.sampler_step_hg(x, None, None, None, None, None, None, None, 2, 1, 4)
    algo = DFoTVideo(_tiny())
    with pytest.raises(RuntimeError, match="CUDA"):
        algo.diffusion_model.model(torch.zeros(1, 8, 4, 8, 8), torch.zeros(1, 8, dtype=torch.long))
    meta, arr, _ = load_case("uvit_pose_vanilla")
    pose = build_product(meta["cfg"])
    cond = pose._window_conditions(torch.from_numpy(arr["conds"]), 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        pose.diffusion_model.model(torch.zeros(1, 4, 3, 32, 32), torch.zeros(1, 4), cond)


@pytest.mark.parametrize("name", ["uvit_pose_vanilla", "uvit_pose_stabilized_interp", "uvit_pose_mean_vanilla",
                                  "uvit_pose_temporal"])
def test_uvit_host_orchestration_with_emulated_kernels(name, monkeypatch):
    """The product's U-ViT3DPose host side (weight packing, channel-last buffer flow, FiLM column offsets, per-window
    pose cache, row maps, K4 tables) driven by CPU restatements of the kernel contracts must reproduce the
    reference rollout.  bf16 operand rounding is emulated too, hence the 2e-2 / 40 dB gates of BASELINE.json."""
    import ops_emulation
    meta, arr, weights = load_case(name)
    cfg = meta["cfg"]
    algo = build_product(cfg)
    sd = {"diffusion_model.model." + k: v for k, v in weights.items()}
    sd["data_mean"], sd["data_std"] = algo.data_mean, algo.data_std
    algo.load_state_dict(sd, strict=True)
    ops_emulation.install(monkeypatch)
    monkeypatch.setattr(ops, "sampler_step_hg", k4_emulation.emulate)
    algo.model_in_dtype = torch.float32
    torch.manual_seed(meta["sampling_seed"])
    algo.diffusion_model.noise_source = lambda shape, device: torch.randn(shape)
    algo.trace = []
    out = algo._predict_videos(torch.from_numpy(arr["xs"]).clone(), cfg["context_frames"], torch.from_numpy(arr["conds"]))
    assert len(algo.trace) == int(arr["n_steps"])
    worst = 0.0
    for i, t in enumerate(algo.trace):
        p = f"step{i:03d}."
        assert np.array_equal(t["levels_from"], arr[p + "levels_from"])
        assert np.array_equal(t["cond_mask"], arr[p + "cond_mask"])
        worst = max(worst, np.abs(t["model_out"].numpy() - arr[p + "model_out"]).max())
    assert worst <= 2e-2, worst
    ref = arr["prediction"][:, cfg["context_frames"]:]
    mse = float(((out.numpy()[:, cfg["context_frames"]:] - ref) ** 2).mean())
    assert 10 * np.log10((ref.max() - ref.min()) ** 2 / max(mse, 1e-30)) >= 40.0


@pytest.mark.parametrize("name", case_names())
def test_planner_end_to_end_on_cpu(name, monkeypatch):
    """Product host logic + K4 contract emulation + oracle backbone (all fp32 on CPU) must reproduce the
    reference rollout: per-step levels bit-exact, tensors <= 2e-5 (coefficients are folded in float64)."""
    meta, arr, weights = load_case(name)
    cfg = meta["cfg"]
    algo = build_product(cfg)
    algo.model_in_dtype = torch.float32
    if "camera_pose_conditioning" in cfg:   # the oracle backbone consumes the reference's dense ray encoding
        from oracle.pose import ray_encoding
        cp = cfg["camera_pose_conditioning"]
        temporal = cfg["tasks"]["prediction"]["history_guidance"]["name"] == "temporal"
        algo._window_conditions = lambda c, nfe, levels_from=None: ray_encoding(
            c.repeat_interleave(nfe, 0), cfg["x_shape"][1], cp["normalize_by"], cp["bound"], cp["type"],
            interp_mask=torch.from_numpy(levels_from == cfg["diffusion"]["timesteps"] - 1) if temporal else None)
    _, backbone = build_oracle(cfg, weights)

    class OracleBackbone(torch.nn.Module):
        def forward(self, x, k, c=None, cm=None, out_dtype=None):
            return backbone(x, k, c, cm)

    algo.diffusion_model.model = OracleBackbone()
    monkeypatch.setattr(ops, "sampler_step_hg", k4_emulation.emulate)
    torch.manual_seed(meta["sampling_seed"])
    algo.diffusion_model.noise_source = lambda shape, device: torch.randn(shape)
    algo.trace = []
    xs = torch.from_numpy(arr["xs"])
    conds = torch.from_numpy(arr["conds"]) if "conds" in arr else None
    out = algo._predict_videos(xs.clone(), cfg["context_frames"], conds)
    assert len(algo.trace) == int(arr["n_steps"])
    for i, t in enumerate(algo.trace):
        p = f"step{i:03d}."
        assert np.array_equal(t["levels_from"], arr[p + "levels_from"]), (name, i)
        assert np.array_equal(t["levels_to"], arr[p + "levels_to"]), (name, i)
        assert (t["cond_mask"] is None) == (p + "cond_mask" not in arr)
        if t["cond_mask"] is not None:
            assert np.array_equal(t["cond_mask"], arr[p + "cond_mask"])
        for k in ["model_in", "model_out"]:
            err = np.abs(t[k].numpy() - arr[p + k]).max()
            assert err <= 2e-5, (name, i, k, err)
    err = np.abs(out.numpy() - arr["prediction"]).max()
    assert err <= 2e-5, (name, err)


# ------------------------------------------------------------------------------------- refinement sampling (fork-only)
def test_refine_scheduling_matrices_bit_exact():
    """`_generate_refine_scheduling_matrix` of the product and of the oracle against the executed reference
    (tests/golden/refine_matrices.json, oracle/make_goldens_refine.py)."""
    from oracle.schedule import refine_scheduling_matrix
    with open(os.path.join(GOLDEN, "refine_matrices.json")) as f:
        recs = json.load(f)
    assert len(recs) >= 5
    for r in recs:
        algo = DFoTVideo(_tiny(**{"diffusion.sampling_timesteps": r["steps"], "max_frames": r["horizon"] + r["padding"]}))
        m = algo._generate_refine_scheduling_matrix(r["horizon"], r["goback_length"], r["n_goback"], r["padding"])
        assert m.dtype == torch.int64 and m.tolist() == r["matrix"]
        o = refine_scheduling_matrix(r["horizon"], r["goback_length"], r["n_goback"], r["padding"], 1000, r["steps"])
        assert o.tolist() == r["matrix"]


def test_q_sample_from_x_k_matches_oracle(monkeypatch):
    """API parity of `q_sample_from_x_k` (discrete_diffusion.py:252-260): the K4 records built on the host (contract
    emulation) against the oracle's restatement, context (-1), pad (999) and ordinary levels mixed; continuous schedule
    (with the discrete cosine schedule ᾱ[-1] = 0 and the context scale is 0/0 in the reference too — quirk Q11)."""
    from oracle.cases import continuous_overrides
    from oracle.diffusion import Diffusion
    monkeypatch.setattr(ops, "sampler_step_hg", k4_emulation.emulate)
    monkeypatch.setattr(ops, "require_cuda", lambda *a, **k: None)
    cfg = _tiny(**continuous_overrides())
    algo = DFoTVideo(cfg)
    g = torch.Generator().manual_seed(0)
    x = torch.randn((2, 4, 4, 8, 8), generator=g)
    noise = torch.randn((2, 4, 4, 8, 8), generator=g)
    cur = torch.tensor([[-1, 499, 499, 999], [-1, 19, 665, 999]])
    nxt = torch.tensor([[-1, 665, 832, 999], [-1, 39, 999, 999]])
    got = algo.diffusion_model.q_sample_from_x_k(x, cur, nxt, noise)
    ref = Diffusion(cfg["diffusion"], None).q_sample_from_x_k(x, cur, nxt, noise)
    assert torch.isfinite(ref).all() and (got - ref).abs().max().item() <= 1e-6
    assert torch.equal(got[:, 0], x[:, 0]) and torch.equal(got[:, 3], x[:, 3])      # context and pad tokens untouched
    # the discrete cosine schedule: NaN on the context token in the reference's arithmetic, here as well
    algo = DFoTVideo(_tiny())
    bad = algo.diffusion_model.q_sample_from_x_k(x, cur, nxt, noise)
    ref = Diffusion(_tiny()["diffusion"], None).q_sample_from_x_k(x, cur, nxt, noise)
    assert torch.isnan(bad[:, 0]).all() and torch.isnan(ref[:, 0]).all() and torch.isfinite(bad[:, 1]).all()


@pytest.mark.parametrize("name", [n for n in case_names() if not n.startswith("uvit")])
def test_dit3d_host_orchestration_with_emulated_kernels(name, monkeypatch):
    """The product's DiT3D host side (weight packing, per-frame modulation columns, RoPE table, QKV / gate epilogue
    arguments, action / label conditioning, K4 tables, refinement walk) driven by CPU restatements of the kernel
    contracts must reproduce every reference rollout.  bf16 operand rounding is emulated, hence the 2e-2 / 40 dB gates."""
    import ops_emulation
    meta, arr, weights = load_case(name)
    cfg = meta["cfg"]
    algo = build_product(cfg)
    sd = {"diffusion_model.model." + k: v for k, v in weights.items()}
    sd["data_mean"], sd["data_std"] = algo.data_mean, algo.data_std
    algo.load_state_dict(sd, strict=True)
    ops_emulation.install(monkeypatch)
    monkeypatch.setattr(ops, "sampler_step_hg", k4_emulation.emulate)
    algo.model_in_dtype = torch.float32
    algo.diffusion_model.model.use_cuda_graph = False
    torch.manual_seed(meta["sampling_seed"])
    algo.diffusion_model.noise_source = lambda shape, device: torch.randn(shape)
    algo.trace = []
    conds = torch.from_numpy(arr["conds"]) if "conds" in arr else None
    out = algo._predict_videos(torch.from_numpy(arr["xs"]).clone(), cfg["context_frames"], conds)
    assert len(algo.trace) == int(arr["n_steps"])
    worst = 0.0
    for i, t in enumerate(algo.trace):
        p = f"step{i:03d}."
        assert np.array_equal(t["levels_from"], arr[p + "levels_from"])
        assert np.array_equal(t["levels_to"], arr[p + "levels_to"])
        worst = max(worst, np.abs(t["model_out"].numpy() - arr[p + "model_out"]).max())
    assert worst <= 2e-2, worst
    n_ctx = cfg["context_frames"]
    ref = arr["prediction"][:, n_ctx:]
    mse = float(((out.numpy()[:, n_ctx:] - ref) ** 2).mean())
    assert 10 * np.log10((ref.max() - ref.min()) ** 2 / max(mse, 1e-30)) >= 40.0


@pytest.mark.parametrize("combo", range(len(MATRIX_COMBOS)))
def test_matrix_attention_combinations_vs_oracle_with_emulated_kernels(combo, monkeypatch):
    """Matrix-attention block types x head groupings x RoPE modes x bias that no golden rollout covers: one forward of the
    product's host side on the kernel-contract emulations against the oracle (which the goldens pin, general in all of these)."""
    import ops_emulation
    model, oracle, x, lv = matrix_combo_model(MATRIX_COMBOS[combo])
    ops_emulation.install(monkeypatch)
    model.use_cuda_graph = False
    want = oracle(x, lv)
    got = model(x, lv)
    assert want.abs().max() > 1e-2
    assert (got - want).abs().max().item() <= 2e-2, (got - want).abs().max().item()


@pytest.mark.parametrize("pos", ["rope_3d", "learned_1d"])
def test_dit3d_splitk_block_loop_equals_plain_loop(pos, monkeypatch):
    """The latency-mode block loop of DiT3D (ops.set_latency_mode; split-K GEMMs at the end of a block half + the gated residual fused into the
    next AdaLN, x never stored) against the plain loop on the kernel-contract emulations, at a width where the k-loops
    really split (hidden 256, MLP x4: 4 splits of fc2) — same arithmetic up to the summation order of the splits."""
    import ops_emulation
    from oracle.cases import _small
    from dfot_b200.algorithms.dfot.backbones.dit.dit3d import DiT3D
    cfg = _small(**{"backbone.hidden_size": 256, "backbone.num_heads": 4, "backbone.spatial_mlp_ratio": 4.0,
                    "backbone.depth": 3, "backbone.pos_emb_type": pos})["backbone"]
    torch.manual_seed(0)
    model = DiT3D(cfg, [4, 8, 8], 4, use_causal_mask=False).eval()
    for prm in model.parameters():                       # AdaLN-Zero / final layer are zero-initialised: redraw
        if prm.abs().sum() == 0:
            torch.nn.init.normal_(prm, std=0.05)
    ops_emulation.install(monkeypatch)
    model.use_cuda_graph = False
    assert ops.splitk_factor(2 * 4 * 16, 256, 1024) == 4 and ops.splitk_factor(128, 256, 256) == 1
    x, lv = torch.randn((2, 4, 4, 8, 8)), torch.randint(0, 1000, (2, 4))
    assert not model._use_splitk(2 * 4 * 16)             # latency mode is off by default
    out_plain = model(x, lv).clone()
    monkeypatch.setattr(ops, "_latency_mode", True)
    assert model._use_splitk(2 * 4 * 16)
    out_split = model(x, lv).clone()
    assert out_plain.abs().max() > 1e-2
    # the fp32 token streams differ in their last bits (summation order of the splits); where that flips the bf16 rounding of
    # a GEMM operand element the outputs move by up to ~1e-3 (seeds 0-5: 0 ... 1.2e-3), so the gate is the flip level
    assert (out_split - out_plain).abs().max().item() <= 2e-3 * max(1.0, out_plain.abs().max().item())


def test_matrix_token_attention_head_dim_is_checked_at_construction():
    """attn2 of a MatrixSelf / MatrixCrossDiTBlock runs num_row_heads heads of embed_row_dim / num_row_heads inside a frame: a
    width the attention kernel does not cover is refused when the model is built, not at the first forward (found by the
    randomised GPU test: two row heads of 32 pass the matrix-attention check — 2 rows x 32 = 64 — but not attn2's)."""
    from oracle.cases import MATRIX_COMBOS, matrix_combo_cfg
    from dfot_b200.algorithms.dfot.backbones.dit.dit3d import DiT3D
    ok = ("full_matrix_attention", "matrix", 2, 1, 2, 64, True, False, False, False, None)          # 2 rows x 32 = 64
    DiT3D(matrix_combo_cfg(ok), [4, 8, 8], 4, use_causal_mask=False)
    for block in ("matrix_self", "matrix_cross"):
        bad = (ok[0], block) + ok[2:]
        with pytest.raises(NotImplementedError, match="token attention head dim 32"):
            DiT3D(matrix_combo_cfg(bad), [4, 8, 8], 4, use_causal_mask=False)
