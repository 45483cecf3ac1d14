"""Randomised host-logic checks (no GPU): the product's DiT3D forward and its sampler driven by the CPU restatements of the
kernel contracts (ops_emulation / k4_emulation) against the oracle, over configurations DRAWN AT RANDOM from the reference's
knobs — backbone variant x position embedding x matrix block / head grouping / RoPE mode / bias x conditioning x window length,
and scheduling matrix x guidance scheme x objective x window plan (sliding context, keyframes + interpolation) x eta x
conditioning.  A configuration on which the ORACLE raises is one the reference rejects as well (checked against the executed
reference for the three classes the generator produces: a condition tensor shorter than a padded tail window, keyframes with
two context frames, explicit `gen_segments` under an autoregressive schedule) and is skipped.  Fixed seeds: the cases are the
same in every run; `python tests/test_fuzz_host.py <seed> <n>` draws more."""
import json
import os
import random
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # (script mode; pytest has conftest.py)

from dfot_b200 import ops  # noqa: E402
from helpers import NoiseBank, build_oracle, build_product  # noqa: E402
from oracle.cases import _small, continuous_overrides  # noqa: E402
import k4_emulation  # noqa: E402
import ops_emulation  # noqa: E402


def _redraw(module):
    for prm in module.parameters():                      # AdaLN-Zero / final layers / biases start at zero
        if prm.abs().sum() == 0:
            torch.nn.init.normal_(prm, std=0.05)


def dit_forward_case(rng: random.Random):
    """(backbone cfg, constructor kwargs, frames, with-mask) or None when the draw is outside what the kernels cover."""
    variant = rng.choice(["full", "factorized_encoder", "factorized_attention", "full_matrix_attention",
                          "factorized_matrix_attention"])
    o = {"backbone.variant": variant, "backbone.depth": rng.choice([1, 2]), "backbone.mlp_ratio": rng.choice([2.0, 4.0]),
         "backbone.spatial_mlp_ratio": rng.choice([None, 2.0])}
    erd = rng.choice([64, 128])
    if "matrix" in variant:
        nrh, nch, n = rng.choice([1, 2]), rng.choice([1, 2]), rng.choice([1, 2])
        multi = rng.choice([False, True])
        flat = (not multi) and rng.choice([False, True])
        d = erd // nrh
        block = rng.choice(["matrix", "matrix_self", "matrix_cross"])
        if (d if multi else n * d) not in (64, 72, 128) or (block != "matrix" and d not in (64, 72, 128)):
            return None                                    # outside what the attention kernel covers (the constructor refuses)
        o.update({"backbone.pos_emb_type": "sinusoidal_2d", "backbone.hidden_size": erd, "backbone.embed_row_dim": erd,
                  "backbone.embed_col_dim": nch * n, "backbone.num_col_heads": nch, "backbone.num_row_heads": nrh,
                  "backbone.num_heads": erd // 64, "backbone.use_temporal_rope": rng.choice([False, True]),
                  "backbone.flatten_matrix_rope": flat, "backbone.matrix_multi_token": multi,
                  "backbone.use_bias": rng.choice([False, True]), "backbone.spatial_mlp_ratio": 2.0,
                  "backbone.matrix_block": block, "backbone.fixed_u": None})
    else:
        o.update({"backbone.hidden_size": erd, "backbone.num_heads": erd // 64,
                  "backbone.pos_emb_type": rng.choice(["rope_3d", "learned_1d", "sinusoidal_1d"] if variant == "full" else
                                                      ["sinusoidal_factorized", "learned_1d", "sinusoidal_1d"])})
    cond = rng.choice([None, "action"])
    if cond:
        o["backbone.external_cond_dropout"] = rng.choice([0.0, 0.1])
    kw = dict(external_cond_type=cond, external_cond_num_classes=None, external_cond_dim=3 if cond else 0)
    return _small(**o)["backbone"], kw, rng.choice([4, 3, 2]), bool(cond) and rng.random() < 0.5


@pytest.mark.parametrize("seed", range(24))
def test_random_dit3d_forward_vs_oracle(seed, monkeypatch):
    from oracle.dit3d import DiT3DOracle
    from dfot_b200.algorithms.dfot.backbones.dit.dit3d import DiT3D
    rng = random.Random(1000 + seed)
    case = None
    while case is None:
        case = dit_forward_case(rng)
    cfg, kw, T, with_mask = case
    torch.manual_seed(seed)
    model = DiT3D(cfg, [4, 8, 8], 4, use_causal_mask=False, **kw).eval()
    _redraw(model)
    model.use_cuda_graph = False
    oracle = DiT3DOracle(cfg, [4, 8, 8], 4, {k: v.detach().clone() for k, v in model.state_dict().items()},
                         external_cond_dim=kw["external_cond_dim"])
    ops_emulation.install(monkeypatch)
    g = torch.Generator().manual_seed(seed)
    x, lv = torch.randn((2, T, 4, 8, 8), generator=g), torch.randint(0, 1000, (2, T), generator=g)
    c = torch.randn((2, T, 3), generator=g) if kw["external_cond_dim"] else None
    m = (torch.rand((2,), generator=g) < 0.5) if with_mask else None
    want, got = oracle(x, lv, c, m), model(x, lv, c, m)
    assert want.abs().max() > 1e-2
    assert (got - want).abs().max().item() <= 2e-2, (cfg, (got - want).abs().max().item())


def uvit_case(rng: random.Random):
    """U-ViT3DPose at golden scale: channels / heads (head dims 64, 128) / blocks per level / mid blocks / resolution / frames /
    pose normalisation drawn at random."""
    from oracle.cases import algorithm_cfg
    heads = rng.choice([1, 2])
    c2 = 64 * heads * rng.choice([1, 2]) if heads == 1 else 128 * rng.choice([1, 2])
    c3 = rng.choice([c2, 2 * c2]) if 2 * c2 // heads <= 128 else c2
    channels = [32, rng.choice([32, 64]), c2, c3]
    updown = [rng.choice([1, 2]), rng.choice([1, 2]), rng.choice([1, 2])]
    res, frames = 32, rng.choice([2, 3, 4])
    return algorithm_cfg(**{
        **continuous_overrides(), "external_cond_type": "action", "external_cond_dim": 16,
        "camera_pose_conditioning": dict(normalize_by=rng.choice(["first", "mean"]), bound=None, type="ray_encoding"),
        "backbone": dict(name="u_vit3d_pose", channels=channels, emb_channels=64, patch_size=2,
                         block_types=["ResBlock", "ResBlock", "TransformerBlock", "TransformerBlock"],
                         block_dropouts=[0.0] * 4, num_updown_blocks=updown, num_mid_blocks=rng.choice([1, 2]),
                         num_heads=heads, pos_emb_type="rope", use_checkpointing=[False] * 4, conditioning=dict(dim=None),
                         external_cond_dropout=0.1, use_fourier_noise_embedding=rng.choice([True, False])),
        "x_shape": [3, res, res], "max_frames": frames, "n_frames": frames, "context_frames": 1,
        "data_mean": [[[0.5]]] * 3, "data_std": [[[0.5]]] * 3, "diffusion.sampling_timesteps": 3}), res, frames


@pytest.mark.parametrize("seed", range(8))
def test_random_uvit3d_pose_forward_vs_oracle(seed, monkeypatch):
    """Cached-pose fast path and dense reference-style conditioning, one row's pose masked, vs the oracle."""
    from oracle.cases import synthetic_poses
    from oracle.pose import ray_encoding
    from oracle.uvit3d_pose import UViT3DPoseOracle
    from dfot_b200.algorithms.dfot import DFoTVideoPose
    cfg, res, frames = uvit_case(random.Random(3000 + seed))
    torch.manual_seed(seed)
    algo = DFoTVideoPose(json.loads(json.dumps(cfg))).eval()
    _redraw(algo)
    model = algo.diffusion_model.model
    oracle = UViT3DPoseOracle(cfg["backbone"], cfg["x_shape"], frames, {k: v.detach().clone() for k, v in model.state_dict().items()})
    ops_emulation.install(monkeypatch)
    model.use_cuda_graph = False
    g = torch.Generator().manual_seed(seed)
    x, levels = torch.randn((2, frames, 3, res, res), generator=g), torch.randn((2, frames), generator=g)
    poses = synthetic_poses(1, frames)
    norm = cfg["camera_pose_conditioning"]["normalize_by"]
    enc = ray_encoding(poses.repeat_interleave(2, 0), res, norm, None, "ray_encoding")
    mask = torch.tensor([True, False])
    want = oracle(x, levels, enc, mask)
    got = model(x, levels, algo._window_conditions(poses, 2), mask)
    assert (got - want).abs().max().item() <= 2e-2, (cfg["backbone"], (got - want).abs().max().item())
    got = model(x, levels, enc, mask)
    assert (got - want).abs().max().item() <= 2e-2, (cfg["backbone"], "dense", (got - want).abs().max().item())


def sampler_case(rng: random.Random):
    o = {"backbone.depth": 1}
    ctx = rng.choice([1, 2])
    o["scheduling_matrix"] = rng.choice(["full_sequence", "autoregressive"])
    o["diffusion.sampling_timesteps"] = rng.choice([3, 4, 5])
    if rng.random() < 0.4:
        o.update(continuous_overrides())
    elif rng.random() < 0.5:
        o["diffusion.objective"] = rng.choice(["pred_v", "pred_noise", "pred_x0"])
    hg = rng.choice(["conditional", "vanilla", "stabilized_vanilla", "fractional", "temporal"])
    scheme = {"conditional": dict(name="conditional", visualize=False),
              "vanilla": dict(name="vanilla", guidance_scale=rng.choice([1.5, 4.0]), visualize=False),
              "stabilized_vanilla": dict(name="stabilized_vanilla", guidance_scale=2.0,
                                         stabilization_level=rng.choice([0.02, 0.1]), visualize=False),
              "fractional": dict(name="fractional", guidance_scale=3.0, freq_scale=0.3, visualize=False),
              "temporal": dict(name="temporal", hist_subsequences=[[0], [1]], hist_weights=[1.5, 1.5],
                               gen_segments=[[0], [1], [0, 1]], visualize=False)}[hg]
    if hg == "temporal":
        ctx = 2
    o["context_frames"], o["tasks.prediction.history_guidance"] = ctx, scheme
    n_frames = rng.choice([4, 4, 6, 7, 9])
    o["n_frames"] = n_frames
    if n_frames > 4:
        if rng.random() < 0.5 and hg != "temporal":
            o["tasks.prediction.keyframe_density"], o["tasks.prediction.sliding_context_len"] = 0.5, ctx
            o["tasks.interpolation.history_guidance"] = rng.choice([dict(name="vanilla", guidance_scale=1.5, visualize=False),
                                                                   dict(name="conditional", visualize=False)])
            o["tasks.interpolation.max_batch_size"] = rng.choice([None, 1, 2])
        else:
            o["tasks.prediction.sliding_context_len"] = rng.choice([ctx, min(ctx + 1, 3)])
    if rng.random() < 0.3:
        o["noise_level"] = "random_uniform"
    if rng.random() < 0.3:
        o["diffusion.ddim_sampling_eta"] = 0.5
    cond = rng.random() < 0.4
    if cond:
        o.update({"external_cond_type": "action", "external_cond_dim": 3,
                  "external_cond_processing": rng.choice([None, "mask_first"]), "backbone.external_cond_dropout": 0.1})
    return _small(**o), n_frames, ctx, rng.choice([1, 2]), cond


def run_sampler_case(case, seed):
    """None when the oracle (hence the reference) rejects the configuration; else (levels equal, per-step max-abs, final
    max-abs with NaNs — pred_noise at the cosine schedule's alpha_bar = 0 — required at the same places, context kept)."""
    cfg, n_frames, ctx, B, cond = case
    torch.manual_seed(seed)
    algo = build_product(json.loads(json.dumps(cfg)))
    _redraw(algo)
    weights = {k[len("diffusion_model.model."):]: v.detach().clone() for k, v in algo.state_dict().items()
               if k.startswith("diffusion_model.model.")}
    bank = NoiseBank(100 + seed)
    oracle, _ = build_oracle(json.loads(json.dumps(cfg)), weights, randn=bank.randn, randn_like=bank.randn_like)
    oracle.trace = []
    g = torch.Generator().manual_seed(seed)
    xs = torch.randn((B, n_frames, 4, 8, 8), generator=g)
    conds = torch.randn((B, n_frames, 3), generator=g) if cond else None
    try:
        with torch.no_grad():
            ref = oracle.predict_videos(xs.clone(), ctx, conds)
    except (IndexError, RuntimeError):
        return None
    bank2 = NoiseBank(100 + seed)
    algo.model_in_dtype = torch.float32
    algo.diffusion_model.model.use_cuda_graph = False
    algo.diffusion_model.noise_source = lambda shape, device: bank2.randn(shape)
    algo.trace = []
    out = algo._predict_videos(xs.clone(), ctx, conds)
    assert len(algo.trace) == len(oracle.trace)
    worst = 0.0
    for t, q in zip(algo.trace, oracle.trace):
        assert np.array_equal(t["levels_from"], q["levels_from"].numpy()) and np.array_equal(t["levels_to"], q["levels_to"].numpy())
        worst = max(worst, (t["model_out"] - q["model_out"]).abs().max().item())
    assert torch.equal(torch.isnan(out), torch.isnan(ref)) and torch.equal(out[:, :ctx], xs[:, :ctx])
    return worst, (torch.nan_to_num(out) - torch.nan_to_num(ref)).abs().max().item()


@pytest.mark.parametrize("seed", range(16))
def test_random_sampler_rollout_vs_oracle(seed, monkeypatch):
    ops_emulation.install(monkeypatch)
    monkeypatch.setattr(ops, "sampler_step_hg", k4_emulation.emulate)
    rng = random.Random(2000 + seed)
    res = None
    while res is None:
        res = run_sampler_case(sampler_case(rng), seed)
    assert res[0] <= 3e-2 and res[1] <= 0.1, res


def pose_sampler_case(rng: random.Random):
    """DFoTVideoPose rollouts: guidance scheme x schedule x pose normalisation / bound x window plan on a small U-ViT3DPose."""
    from oracle.cases import algorithm_cfg
    ctx = rng.choice([1, 2])
    hg = rng.choice(["conditional", "vanilla", "stabilized_vanilla", "fractional", "temporal"])
    scheme = {"conditional": dict(name="conditional", visualize=False),
              "vanilla": dict(name="vanilla", guidance_scale=2.0, visualize=False),
              "stabilized_vanilla": dict(name="stabilized_vanilla", guidance_scale=2.0, stabilization_level=0.02, visualize=False),
              "fractional": dict(name="fractional", guidance_scale=3.0, freq_scale=0.3, visualize=False),
              "temporal": dict(name="temporal", hist_subsequences=[[0], [1]], hist_weights=[1.5, 1.5],
                               gen_segments=[[0], [1], [0, 1]], visualize=False)}[hg]
    if hg == "temporal":
        ctx = 2
    n_frames = rng.choice([4, 4, 6, 7, 9])
    o = {**continuous_overrides(), "external_cond_type": "action", "external_cond_dim": 16,
         "camera_pose_conditioning": dict(normalize_by=rng.choice(["first", "mean"]), bound=rng.choice([None, 10.0]),
                                          type="ray_encoding"),
         "backbone": dict(name="u_vit3d_pose", channels=[32, 32, 64, 128], emb_channels=64, patch_size=2,
                          block_types=["ResBlock", "ResBlock", "TransformerBlock", "TransformerBlock"],
                          block_dropouts=[0.0] * 4, num_updown_blocks=[1, 1, 1], num_mid_blocks=1, num_heads=1,
                          pos_emb_type="rope", use_checkpointing=[False] * 4, conditioning=dict(dim=None),
                          external_cond_dropout=0.1, use_fourier_noise_embedding=True),
         "x_shape": [3, 32, 32], "max_frames": 4, "n_frames": n_frames, "context_frames": ctx,
         "data_mean": [[[0.5]]] * 3, "data_std": [[[0.5]]] * 3, "diffusion.sampling_timesteps": rng.choice([2, 3]),
         "scheduling_matrix": rng.choice(["full_sequence", "autoregressive"]), "tasks.prediction.history_guidance": scheme}
    if n_frames > 4:
        if rng.random() < 0.5 and hg != "temporal":
            o["tasks.prediction.keyframe_density"], o["tasks.prediction.sliding_context_len"] = 0.5, ctx
            o["tasks.interpolation.history_guidance"] = rng.choice([dict(name="vanilla", guidance_scale=1.5, visualize=False),
                                                                   dict(name="conditional", visualize=False)])
            o["tasks.interpolation.max_batch_size"] = rng.choice([None, 1, 2])
            o["tasks.interpolation.enabled"] = rng.choice([False, True])
        else:
            o["tasks.prediction.sliding_context_len"] = rng.choice([ctx, min(ctx + 1, 3)])
    return algorithm_cfg(**o), n_frames, ctx, rng.choice([1, 2])


@pytest.mark.parametrize("seed", range(6))
def test_random_pose_sampler_rollout_vs_oracle(seed, monkeypatch):
    from oracle.cases import synthetic_poses
    ops_emulation.install(monkeypatch)
    monkeypatch.setattr(ops, "sampler_step_hg", k4_emulation.emulate)
    rng = random.Random(4000 + seed)
    while True:
        cfg, n_frames, ctx, B = pose_sampler_case(rng)
        torch.manual_seed(seed)
        algo = build_product(json.loads(json.dumps(cfg)))
        _redraw(algo)
        weights = {k[len("diffusion_model.model."):]: v.detach().clone() for k, v in algo.state_dict().items()
                   if k.startswith("diffusion_model.model.")}
        bank = NoiseBank(100 + seed)
        oracle, _ = build_oracle(json.loads(json.dumps(cfg)), weights, randn=bank.randn, randn_like=bank.randn_like)
        oracle.trace = []
        g = torch.Generator().manual_seed(seed)
        xs, conds = torch.randn((B, n_frames, 3, 32, 32), generator=g), synthetic_poses(B, n_frames)
        try:
            with torch.no_grad():
                ref = oracle.predict_videos(xs.clone(), ctx, conds)
            break
        except (IndexError, RuntimeError):          # rejected by the reference as well (see the module docstring): draw again
            continue
    bank2 = NoiseBank(100 + seed)
    algo.model_in_dtype = torch.float32
    algo.diffusion_model.model.use_cuda_graph = False
    algo.diffusion_model.noise_source = lambda shape, device: bank2.randn(shape)
    algo.trace = []
    out = algo._predict_videos(xs.clone(), ctx, conds)
    assert len(algo.trace) == len(oracle.trace)
    for t, q in zip(algo.trace, oracle.trace):
        assert np.array_equal(t["levels_from"], q["levels_from"].numpy())
        # (weights re-drawn at std 0.05 and the steps of a rollout compound the emulated bf16 rounding: a host-logic slip
        # — wrong window, branch weight, pose row — shows up as an O(1) error, the gate only has to sit well below that)
        assert (t["model_out"] - q["model_out"]).abs().max().item() <= 6e-2
    assert (out - ref).abs().max().item() <= 0.15 and torch.equal(out[:, :ctx], xs[:, :ctx])


if __name__ == "__main__":          # python tests/test_fuzz_host.py <seed> <n>: more sampler draws than the test runs
    class _MP:
        def setattr(self, obj, name, val, raising=True):
            setattr(obj, name, val)
    ops_emulation.install(_MP())
    ops.sampler_step_hg = k4_emulation.emulate
    rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
        case = sampler_case(rng)
        print(i, case[0]["scheduling_matrix"], case[0]["tasks"]["prediction"]["history_guidance"]["name"], case[1:],
              run_sampler_case(case, i))
