"""-m gpu: the CUDA sampling path (through the C ABI) against the oracle and the committed goldens.

Gates (BASELINE.json north_star): scheduling / mask / level integers bit-exact; per-step denoiser output
max-abs error <= 2e-2 (bf16 kernels vs fp32 reference); final sample PSNR >= 40 dB vs the reference rollout
(data_range = value range of the reference rollout, non-context frames)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from dfot_b200.algorithms.dfot import DFoTVideo  # noqa: E402
from helpers import MATRIX_COMBOS, NoiseBank, build_oracle, build_product, case_names, load_case, matrix_combo_model  # noqa: E402
from oracle.cases import algorithm_cfg, continuous_overrides  # noqa: E402

DEV = "cuda"
STEP_TOL = 2e-2
PSNR_MIN = 40.0


def psnr(pred, ref, n_ctx):
    p, r = pred[:, n_ctx:].double(), ref[:, n_ctx:].double()
    rng = (r.max() - r.min()).item()
    mse = ((p - r) ** 2).mean().item()
    return 10 * math.log10(rng * rng / max(mse, 1e-30))


def load_product(cfg, weights):
    algo = build_product(cfg)
    sd = {"diffusion_model.model." + k: v for k, v in weights.items()}
    sd["data_mean"], sd["data_std"] = algo.data_mean, algo.data_std
    algo.load_state_dict(sd, strict=True)
    return algo.to(DEV).eval()


@pytest.mark.parametrize("name", case_names())
def test_golden_case_on_gpu(name):
    """GPU rollout vs the fixture produced by executing the reference (same weights, inputs and noise stream)."""
    meta, arr, weights = load_case(name)
    cfg = meta["cfg"]
    algo = load_product(cfg, weights)
    torch.manual_seed(meta["sampling_seed"])
    algo.diffusion_model.noise_source = lambda shape, device: torch.randn(shape).to(device)
    algo.trace = []
    xs = torch.from_numpy(arr["xs"]).to(DEV)
    conds = torch.from_numpy(arr["conds"]).to(DEV) if "conds" in arr else None
    out = algo._predict_videos(xs.clone(), cfg["context_frames"], conds).cpu()
    assert len(algo.trace) == int(arr["n_steps"])
    worst = 0.0
    for i, t in enumerate(algo.trace):
        p = f"step{i:03d}."
        assert np.array_equal(t["levels_from"], arr[p + "levels_from"])
        assert np.array_equal(t["levels_to"], arr[p + "levels_to"])
        if t["cond_mask"] is not None:
            assert np.array_equal(t["cond_mask"], arr[p + "cond_mask"])
        worst = max(worst, np.abs(t["model_out"].cpu().numpy() - arr[p + "model_out"]).max())
    assert worst <= STEP_TOL, f"per-step denoiser output max-abs error {worst}"
    ref = torch.from_numpy(arr["prediction"])
    assert psnr(out, ref, cfg["context_frames"]) >= PSNR_MIN
    assert torch.equal(out[:, :cfg["context_frames"]], ref[:, :cfg["context_frames"]])  # context untouched


def tiny_cfg1(**over):
    """BASELINE.json configs[0]: tiny DFoT DiT3D (4 layers, hidden 256), 8-frame 16x16 latents, 10 steps, 1 context."""
    base = {"backbone.hidden_size": 256, "backbone.depth": 4, "backbone.num_heads": 4, "backbone.spatial_mlp_ratio": 4.0,
            "x_shape": [4, 16, 16], "max_frames": 8, "n_frames": 8, "context_frames": 1,
            "diffusion.sampling_timesteps": 10,
            "tasks.prediction.history_guidance": dict(name="vanilla", guidance_scale=4.0, visualize=False)}
    base.update(over)
    return algorithm_cfg(**base)


def random_weights(cfg, seed):
    """Reference-shaped random init (zero-initialised outputs re-drawn N(0, 0.02) like the goldens)."""
    torch.manual_seed(seed)
    algo = DFoTVideo(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for _, p in algo.named_parameters():
            if bool((p == 0).all()):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return algo


@pytest.mark.parametrize("variant", ["vanilla", "stabilized_sliding", "continuous_action", "nomlp_conditional"])
def test_cfg1_vs_oracle(variant):
    over = {}
    n_frames, batch = 8, 2
    if variant == "stabilized_sliding":
        over = {"n_frames": 14, "tasks.prediction.history_guidance":
                dict(name="stabilized_vanilla", guidance_scale=4.0, stabilization_level=0.02, visualize=False)}
        n_frames, batch = 14, 1
    elif variant == "continuous_action":
        over = {**continuous_overrides(), "external_cond_type": "action", "external_cond_dim": 3,
                "external_cond_processing": "mask_first", "backbone.external_cond_dropout": 0.1}
    elif variant == "nomlp_conditional":
        over = {"backbone.spatial_mlp_ratio": None,
                "tasks.prediction.history_guidance": dict(name="conditional", visualize=False)}
    cfg = tiny_cfg1(**over)
    algo = random_weights(cfg, 0)
    weights = {k[len("diffusion_model.model."):]: v.detach().clone() for k, v in algo.state_dict().items()
               if k.startswith("diffusion_model.model.")}
    g = torch.Generator().manual_seed(123)
    xs = torch.randn((batch, n_frames, 4, 16, 16), generator=g)
    conds = torch.randn((batch, n_frames, 3), generator=g) if cfg["external_cond_dim"] else None
    # oracle on the host CPU
    bank = NoiseBank(7)
    oracle, _ = build_oracle(cfg, weights, randn=bank.randn, randn_like=bank.randn_like)
    oracle.trace = []
    ref = oracle.predict_videos(xs.clone(), 1, conds)
    # product on the GPU with the same noise stream
    bank2 = NoiseBank(7)
    algo = algo.to(DEV).eval()
    algo.diffusion_model.noise_source = lambda shape, device: bank2.randn(shape).to(device)
    algo.trace = []
    out = algo._predict_videos(xs.to(DEV), 1, None if conds is None else conds.to(DEV)).cpu()
    assert len(algo.trace) == len(oracle.trace)
    worst = 0.0
    for t, o in zip(algo.trace, oracle.trace):
        assert np.array_equal(t["levels_from"], o["levels_from"].numpy())
        assert np.array_equal(t["levels_to"], o["levels_to"].numpy())
        assert np.array_equal(t["context_mask"], o["context_mask"].numpy())
        worst = max(worst, (t["model_out"].cpu() - o["model_out"]).abs().max().item())
    assert worst <= STEP_TOL, f"per-step denoiser output max-abs error {worst}"
    assert psnr(out, ref, 1) >= PSNR_MIN


def test_api_level_calls():
    """sample_step / q_sample / HistoryGuidance manager — the reference's per-step API — agree with the oracle."""
    from oracle import history_guidance as ohg
    from oracle.diffusion import Diffusion
    from dfot_b200.algorithms.dfot.history_guidance import HistoryGuidance
    cfg = tiny_cfg1()
    algo = random_weights(cfg, 3).to(DEV).eval()
    dm = algo.diffusion_model
    g = torch.Generator().manual_seed(9)
    x = torch.randn((2, 8, 4, 16, 16), generator=g)
    k = torch.randint(0, 1000, (2, 8), generator=g)
    noise = torch.randn(x.shape, generator=g)
    od = Diffusion(cfg["diffusion"], None)
    assert (dm.q_sample(x.to(DEV), k.to(DEV), noise.to(DEV)).cpu() - od.q_sample(x, k, noise)).abs().max() < 1e-5
    mask = torch.tensor([[1, 2, 0, 0, 0, 0, -1, -1]] * 2)
    frm = torch.tensor([[-1, -1, 599, 599, 599, 599, 999, 999]] * 2)
    to = torch.tensor([[-1, -1, 499, 499, 499, 499, 999, 999]] * 2)
    scheme_cfg = dict(name="stabilized_vanilla", guidance_scale=3.0, stabilization_level=0.02)
    bank, bank2 = NoiseBank(11), NoiseBank(11)
    od = Diffusion(cfg["diffusion"], None, bank.randn_like)
    sch = ohg.scheme_from_config(scheme_cfg, 1000)
    tab = ohg.branch_table(sch, mask[0])
    xr, f, t, cm, excl = ohg.full_prepare(sch, tab, mask, x, frm, to, od.q_sample, False, bank.randn_like)
    dm.noise_source = lambda shape, device: bank2.randn(shape).to(device)
    hgd = HistoryGuidance.from_config(scheme_cfg, timesteps=1000)
    with hgd(mask.to(DEV)) as mgr:
        assert mgr.nfe == tab.nfe
        px, pf, pt, pcm = mgr.prepare(x.to(DEV), frm.to(DEV), to.to(DEV), dm.q_sample, False)
        assert torch.equal(pf.cpu(), f) and torch.equal(pt.cpu(), t) and torch.equal(pcm.cpu(), cm)
        assert (px.cpu() - xr).abs().max() < 1e-5
        comp = mgr.compose(px)
        assert (comp.cpu() - ohg.full_compose(tab, excl, xr)).abs().max() < 1e-4


# ------------------------------------------------------------------ U-ViT3DPose (RE10K backbone)
def uvit_cfg(channels, heads, res, frames, updown=(1, 1, 2), mid=2, emb=64, **over):
    base = {**continuous_overrides(), "external_cond_type": "action", "external_cond_dim": 16,
            "camera_pose_conditioning": dict(normalize_by="first", bound=None, type="ray_encoding"),
            "backbone": dict(name="u_vit3d_pose", channels=list(channels), emb_channels=emb, patch_size=2,
                             block_types=["ResBlock", "ResBlock", "TransformerBlock", "TransformerBlock"],
                             block_dropouts=[0.0] * 4, num_updown_blocks=list(updown), num_mid_blocks=mid,
                             num_heads=heads, pos_emb_type="rope", use_checkpointing=[False] * 4,
                             conditioning=dict(dim=None), external_cond_dropout=0.1, use_fourier_noise_embedding=True),
            "x_shape": [3, res, res], "max_frames": frames, "n_frames": frames, "context_frames": 1,
            "data_mean": [[[0.5]]] * 3, "data_std": [[[0.5]]] * 3, "diffusion.sampling_timesteps": 3,
            "tasks.prediction.history_guidance": dict(name="vanilla", guidance_scale=2.0, visualize=False)}
    base.update(over)
    return algorithm_cfg(**base)


def random_pose_algo(cfg, seed):
    from dfot_b200.algorithms.dfot import DFoTVideoPose
    torch.manual_seed(seed)
    algo = DFoTVideoPose(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for _, p in algo.named_parameters():
            if bool((p == 0).all()):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return algo


@pytest.mark.parametrize("channels,heads,res,frames", [((32, 32, 64, 128), 1, 32, 4), ((64, 128, 128, 256), 2, 64, 3)])
def test_uvit3d_pose_forward_vs_oracle(channels, heads, res, frames):
    """One backbone forward (both the dense reference-style conditioning and the cached fast path, with one row's pose
    masked) against the fp32 oracle on identical weights."""
    from oracle.cases import synthetic_poses
    from oracle.pose import ray_encoding
    from oracle.uvit3d_pose import UViT3DPoseOracle
    cfg = uvit_cfg(channels, heads, res, frames)
    algo = random_pose_algo(cfg, 5)
    model = algo.diffusion_model.model
    weights = {k: v.detach().clone() for k, v in model.state_dict().items()}
    oracle = UViT3DPoseOracle(cfg["backbone"], cfg["x_shape"], frames, weights)
    g = torch.Generator().manual_seed(77)
    B, nfe = 2, 2
    R = B * nfe
    x = torch.randn((R, frames, 3, res, res), generator=g)
    levels = torch.randn((R, frames), generator=g)
    poses = synthetic_poses(B, frames)
    enc = ray_encoding(poses.repeat_interleave(nfe, 0), res, "first", None, "ray_encoding")
    mask = torch.tensor([True, False] * B)
    ref = oracle(x, levels, enc, mask)
    algo = algo.to(DEV).eval()
    model = algo.diffusion_model.model
    cond = algo._window_conditions(poses.to(DEV), nfe)
    for attempt in range(3):   # eager, graph capture, graph replay
        out = model(x.to(DEV), levels.to(DEV), cond, mask.to(DEV)).cpu()
        err = (out - ref).abs().max().item()
        assert err <= STEP_TOL, f"fast path (call {attempt}): max-abs error {err} (ref max {ref.abs().max().item()})"
    out = model(x.to(DEV), levels.to(DEV), enc.to(DEV), mask.to(DEV)).cpu()
    err = (out - ref).abs().max().item()
    assert err <= STEP_TOL, f"dense path: max-abs error {err}"


def test_uvit3d_pose_rollout_vs_oracle():
    """Medium U-ViT3DPose (head dims 64/128, 64x64 frames) vanilla-HG rollout vs the oracle with a shared noise stream."""
    from oracle.cases import synthetic_poses
    cfg = uvit_cfg((64, 128, 128, 256), 2, 64, 4)
    algo = random_pose_algo(cfg, 2)
    weights = {k[len("diffusion_model.model."):]: v.detach().clone() for k, v in algo.state_dict().items()
               if k.startswith("diffusion_model.model.")}
    g = torch.Generator().manual_seed(123)
    xs = torch.randn((2, 4, 3, 64, 64), generator=g)
    conds = synthetic_poses(2, 4)
    bank = NoiseBank(17)
    oracle, _ = build_oracle(cfg, weights, randn=bank.randn, randn_like=bank.randn_like)
    oracle.trace = []
    ref = oracle.predict_videos(xs.clone(), 1, conds)
    bank2 = NoiseBank(17)
    algo = algo.to(DEV).eval()
    algo.diffusion_model.noise_source = lambda shape, device: bank2.randn(shape).to(device)
    algo.trace = []
    out = algo._predict_videos(xs.to(DEV), 1, conds.to(DEV)).cpu()
    assert len(algo.trace) == len(oracle.trace)
    worst = 0.0
    for t, o in zip(algo.trace, oracle.trace):
        assert np.array_equal(t["levels_from"], o["levels_from"].numpy())
        assert np.array_equal(t["levels_to"], o["levels_to"].numpy())
        worst = max(worst, (t["model_out"].cpu() - o["model_out"]).abs().max().item())
    assert worst <= STEP_TOL, f"per-step denoiser output max-abs error {worst}"
    assert psnr(out, ref, 1) >= PSNR_MIN


def _rollout_vs_oracle(cfg, algo, xs, conds, n_ctx, seed=29):
    weights = {k[len("diffusion_model.model."):]: v.detach().clone() for k, v in algo.state_dict().items()
               if k.startswith("diffusion_model.model.")}
    bank = NoiseBank(seed)
    oracle, _ = build_oracle(cfg, weights, randn=bank.randn, randn_like=bank.randn_like)
    oracle.trace = []
    ref = oracle.predict_videos(xs.clone(), n_ctx, conds)
    bank2 = NoiseBank(seed)
    algo = algo.to(DEV).eval()
    algo.diffusion_model.noise_source = lambda shape, device: bank2.randn(shape).to(device)
    algo.trace = []
    out = algo._predict_videos(xs.to(DEV), n_ctx, None if conds is None else conds.to(DEV)).cpu()
    assert len(algo.trace) == len(oracle.trace)
    worst = 0.0
    for t, o in zip(algo.trace, oracle.trace):
        assert np.array_equal(t["levels_from"], o["levels_from"].numpy())
        assert np.array_equal(t["levels_to"], o["levels_to"].numpy())
        worst = max(worst, (t["model_out"].cpu() - o["model_out"]).abs().max().item())
    assert worst <= STEP_TOL, f"per-step denoiser output max-abs error {worst}"
    assert psnr(out, ref, n_ctx) >= PSNR_MIN
    return len(algo.trace)


def test_cfg5_shaped_long_context_dit_vs_oracle():
    """BASELINE configs[4] shape family (DMLab): DiT-B width (768, 12 heads of 64), latents [32,8,8] with patch 2
    (16 tokens per frame), 36-frame context window, continuous diffusion, action conditioning with mask_first,
    vanilla history guidance — depth cut to 2 so that the CPU oracle finishes in seconds."""
    over = {**continuous_overrides(), "backbone.hidden_size": 768, "backbone.depth": 2, "backbone.num_heads": 12,
            "backbone.spatial_mlp_ratio": 4.0, "x_shape": [32, 8, 8], "max_frames": 36, "n_frames": 36, "context_frames": 4,
            "external_cond_type": "action", "external_cond_dim": 3, "external_cond_processing": "mask_first",
            "backbone.external_cond_dropout": 0.1, "diffusion.sampling_timesteps": 4,
            "tasks.prediction.history_guidance": dict(name="vanilla", guidance_scale=2.0, visualize=False)}
    cfg = algorithm_cfg(**over)
    algo = random_weights(cfg, 4)
    g = torch.Generator().manual_seed(5)
    xs = torch.randn((2, 36, 32, 8, 8), generator=g)
    conds = torch.randn((2, 36, 3), generator=g)
    assert _rollout_vs_oracle(cfg, algo, xs, conds, 4) == 4


def test_cfg4_shaped_long_rollout_uvit_vs_oracle():
    """BASELINE configs[3] structure at reduced size: single image -> 25 frames with a 4-frame U-ViT3DPose window:
    7 keyframes by two sliding windows under stabilized_vanilla guidance, then rounds of vanilla-guided interpolation
    in batches of 4 chunks (ragged last batch: 6 = 4 + 2), every window with its own camera-pose cache."""
    from oracle.cases import synthetic_poses
    cfg = uvit_cfg((32, 32, 64, 128), 1, 32, 4, **{
        "n_frames": 25, "tasks.prediction.keyframe_density": 0.28, "tasks.prediction.sliding_context_len": 1,
        "diffusion.sampling_timesteps": 2,
        "tasks.prediction.history_guidance": dict(name="stabilized_vanilla", guidance_scale=2.0,
                                                  stabilization_level=0.02, visualize=False),
        "tasks.interpolation.history_guidance": dict(name="vanilla", guidance_scale=1.5, visualize=False),
        "tasks.interpolation.max_batch_size": 4})
    algo = random_pose_algo(cfg, 6)
    g = torch.Generator().manual_seed(8)
    xs = torch.randn((1, 25, 3, 32, 32), generator=g)
    conds = synthetic_poses(1, 25)
    steps = _rollout_vs_oracle(cfg, algo, xs, conds, 1)
    assert steps > 10      # several windows x 2 steps


# ------------------------------------------------------------------ batch invariance / lockstep rounds (multi-GPU contract)
def test_uvit_forward_rows_do_not_depend_on_the_batch():
    """A forward-row's output is bit-identical whether the row is forwarded alone or inside a batch: GEMM / conv tiles,
    attention items and the fixed-point GroupNorm statistics never mix rows.  This is what lets the multi-GPU path deal
    rows over ranks (RowShard) and still reproduce the single-GPU rollout exactly."""
    from oracle.cases import synthetic_poses
    cfg = uvit_cfg((64, 128, 128, 256), 2, 64, 3)
    algo = random_pose_algo(cfg, 5).to(DEV).eval()
    model = algo.diffusion_model.model
    g = torch.Generator().manual_seed(1)
    x = torch.randn((4, 3, 3, 64, 64), generator=g).to(DEV)
    levels = torch.randn((4, 3), generator=g).to(DEV)
    cond = algo._window_conditions(synthetic_poses(2, 3).to(DEV), 2)
    mask = torch.tensor([True, False, True, False], device=DEV)
    full = model(x, levels, cond, mask).clone()
    again = model(x, levels, cond, mask).clone()        # second call = CUDA-graph capture / replay of the same signature
    assert torch.equal(full, again)
    for i in range(4):
        rows = torch.tensor([i])
        one = model(x[i:i + 1].contiguous(), levels[i:i + 1].contiguous(), cond.index_select(0, rows), mask[i:i + 1])
        assert torch.equal(one[0], full[i]), f"row {i}: max diff {(one[0] - full[i]).abs().max().item()}"


def test_dit_forward_rows_do_not_depend_on_the_batch():
    cfg = tiny_cfg1()
    algo = random_weights(cfg, 2).to(DEV).eval()
    model = algo.diffusion_model.model
    g = torch.Generator().manual_seed(4)
    x = torch.randn((4, 8, 4, 16, 16), generator=g).to(DEV)
    k = torch.randint(0, 1000, (4, 8), generator=g).to(DEV)
    full = model(x, k).clone()
    for i in range(4):
        one = model(x[i:i + 1].contiguous(), k[i:i + 1].contiguous())
        assert torch.equal(one[0], full[i]), f"row {i}: max diff {(one[0] - full[i]).abs().max().item()}"


@pytest.mark.parametrize("combo", range(len(MATRIX_COMBOS)))
def test_matrix_attention_combinations_vs_oracle(combo):
    """tests/test_host_logic.py's matrix-attention combinations (block type x head grouping x RoPE mode x bias) on the
    kernels: one forward against the CPU oracle, eager and through the captured graph."""
    model, oracle, x, lv = matrix_combo_model(MATRIX_COMBOS[combo])
    model = model.to(DEV)
    want = oracle(x, lv)
    for _ in range(3):                                   # eager, capture, replay
        got = model(x.to(DEV), lv.to(DEV)).float().cpu()
        assert (got - want).abs().max().item() <= 2e-2, (got - want).abs().max().item()


@pytest.mark.parametrize("seed", range(24))
def test_random_dit3d_forward_on_gpu(seed):
    """tests/test_fuzz_host.py's random DiT3D configurations (variant x positions x matrix block / head grouping / RoPE mode /
    bias x conditioning x window length) on the kernels: eager forward and captured graph vs the CPU oracle."""
    import random
    from oracle.dit3d import DiT3DOracle
    from dfot_b200.algorithms.dfot.backbones.dit.dit3d import DiT3D
    from test_fuzz_host import _redraw, dit_forward_case
    rng = random.Random(1000 + seed)
    case = None
    while case is None:
        case = dit_forward_case(rng)
    cfg, kw, T, with_mask = case
    torch.manual_seed(seed)
    model = DiT3D(cfg, [4, 8, 8], 4, use_causal_mask=False, **kw).eval()
    _redraw(model)
    oracle = DiT3DOracle(cfg, [4, 8, 8], 4, {k: v.detach().clone() for k, v in model.state_dict().items()},
                         external_cond_dim=kw["external_cond_dim"])
    g = torch.Generator().manual_seed(seed)
    x, lv = torch.randn((2, T, 4, 8, 8), generator=g), torch.randint(0, 1000, (2, T), generator=g)
    c = torch.randn((2, T, 3), generator=g) if kw["external_cond_dim"] else None
    m = (torch.rand((2,), generator=g) < 0.5) if with_mask else None
    want = oracle(x, lv, c, m)
    model = model.to(DEV)
    dev = lambda t: None if t is None else t.to(DEV)
    for _ in range(3):                                   # eager, capture, replay
        got = model(dev(x), dev(lv), dev(c), dev(m)).float().cpu()
        assert (got - want).abs().max().item() <= 2e-2, (cfg, (got - want).abs().max().item())


def test_lockstep_rounds_reproduce_the_sequential_rollout_on_gpu():
    """BASELINE configs[3] structure (keyframe windows + interpolation rounds): with a row shard the chunk batches of a
    round advance in lockstep, every batch drawing from its own position of torch's CUDA generator stream.  On one GPU
    (a world of 1) that must give the sequential rollout bit for bit and leave the generator where the sequential run
    leaves it."""
    from dfot_b200 import distributed as D
    from oracle.cases import synthetic_poses
    cfg = uvit_cfg((32, 32, 64, 128), 1, 32, 4, **{
        "n_frames": 25, "tasks.prediction.keyframe_density": 0.28, "tasks.prediction.sliding_context_len": 1,
        "diffusion.sampling_timesteps": 3,
        "tasks.prediction.history_guidance": dict(name="stabilized_vanilla", guidance_scale=2.0,
                                                  stabilization_level=0.02, visualize=False),
        "tasks.interpolation.history_guidance": dict(name="vanilla", guidance_scale=1.5, visualize=False),
        "tasks.interpolation.max_batch_size": 2})
    algo = random_pose_algo(cfg, 6).to(DEV).eval()
    g = torch.Generator().manual_seed(8)
    xs = torch.randn((1, 25, 3, 32, 32), generator=g).to(DEV)
    conds = synthetic_poses(1, 25).to(DEV)
    torch.manual_seed(11)
    seq = algo._predict_videos(xs, 1, conds).clone()
    state_seq = torch.cuda.get_rng_state()
    rows_seq = algo.nfe_rows
    torch.manual_seed(11)
    algo.row_shard = D.RowShard(world=1, rank=0)
    try:
        lock = algo._predict_videos(xs, 1, conds).clone()
    finally:
        algo.row_shard = None
    assert torch.equal(torch.cuda.get_rng_state(), state_seq)
    assert algo.nfe_rows == 2 * rows_seq
    assert torch.equal(lock, seq), f"max diff {(lock - seq).abs().max().item()}"
