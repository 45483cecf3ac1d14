"""-m gpu: model-level parity at the two FULL-SIZE benchmarked configurations (BASELINE.json configs[1] and [2]).

The models are the ones bench.py builds (`bench.k600_cfg()` DiT3D-XL: 28 blocks, hidden 1152, 16 heads of 72;
`bench.re10k_cfg()` U-ViT3DPose: 128/256/576/1152 channels, 3/3/6 + 20 mid blocks = 44, N = 8192 tokens at level 2),
random-init with the zero-initialised outputs re-drawn, at B = 1.  One backbone forward-row and a 2-step rollout
(vanilla history guidance for RE10K, the configuration's conditional guidance for K600) run on the GPU through the C ABI
and on the CPU oracle (torch fp32; oracle/uvit3d_pose.py, oracle/dit3d.py) with a shared NoiseBank.

Gates (BASELINE.json north_star): integer levels bit-exact, per-step denoiser output max-abs <= 2e-2, PSNR >= 40 dB.
Reference: algorithms/dfot/backbones/u_vit/u_vit3d.py:199-282, backbones/dit/dit_base.py:310-425.
"""
import json
import math
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from helpers import NoiseBank, build_oracle  # noqa: E402

DEV = "cuda"
STEP_TOL = 2e-2
PSNR_MIN = 40.0


def _psnr(pred, ref, n_ctx):
    p, r = pred[:, n_ctx:].double(), ref[:, n_ctx:].double()
    rng = (r.max() - r.min()).item()
    return 10 * math.log10(rng * rng / max(((p - r) ** 2).mean().item(), 1e-30))


def _weights(algo):
    return {k[len("diffusion_model.model."):]: v.detach().clone() for k, v in algo.state_dict().items()
            if k.startswith("diffusion_model.model.")}


def _rollout(cfg, xs, conds, n_ctx, seed):
    torch.set_num_threads(os.cpu_count() or 1)
    algo = bench.make_weights(cfg, 0)
    bank = NoiseBank(seed)
    oracle, _ = build_oracle(json.loads(json.dumps(cfg)), _weights(algo), randn=bank.randn, randn_like=bank.randn_like)
    oracle.trace = []
    with torch.no_grad():
        ref = oracle.predict_videos(xs.clone(), n_ctx, conds)
    bank2 = NoiseBank(seed)
    algo = algo.to(DEV).eval()
    algo.diffusion_model.noise_source = lambda shape, device: bank2.randn(shape).to(device)
    algo.trace = []
    out = algo._predict_videos(xs.to(DEV), n_ctx, None if conds is None else conds.to(DEV)).cpu()
    assert len(algo.trace) == len(oracle.trace) == cfg["diffusion"]["sampling_timesteps"]
    errs, mags = [], []
    for t, o in zip(algo.trace, oracle.trace):
        assert np.array_equal(t["levels_from"], o["levels_from"].numpy())
        assert np.array_equal(t["levels_to"], o["levels_to"].numpy())
        assert np.array_equal(t["context_mask"], o["context_mask"].numpy())
        errs.append((t["model_out"].cpu() - o["model_out"]).abs().max().item())
        mags.append(o["model_out"].abs().max().item())
    db = _psnr(out, ref, n_ctx)
    print(f"full-size parity: per-step max-abs {errs} (|ref| max {mags}), PSNR {db:.1f} dB")
    assert max(errs) <= STEP_TOL, f"per-step denoiser output max-abs error {errs}"
    assert db >= PSNR_MIN
    assert torch.equal(out[:, :n_ctx], xs[:, :n_ctx])
    return errs, db


def test_re10k_fullsize_rollout_vs_oracle():
    """BASELINE configs[2]: dfot_video_pose, U-ViT3DPose 44 blocks, 8 frames 256x256, vanilla HG 4.0 (2 branch rows per
    step): 2 DDIM steps = 4 forward-rows of the full-size backbone on the oracle."""
    cfg = bench.re10k_cfg(sampling_timesteps=2)
    g = torch.Generator().manual_seed(123)
    xs = torch.rand((1, 8, 3, 256, 256), generator=g)
    mean = torch.tensor(cfg["data_mean"]).reshape(1, 1, 3, 1, 1)
    std = torch.tensor(cfg["data_std"]).reshape(1, 1, 3, 1, 1)
    xs = (xs - mean) / std
    _rollout(cfg, xs, bench.synthetic_poses(1, 8), 1, seed=41)


@pytest.mark.parametrize("mlp", [4.0, None])
def test_k600_fullsize_rollout_vs_oracle(mlp):
    """BASELINE configs[1]: DiT3D-XL (28 x 1152, 16 heads of 72, patch 1) on 16x16x16 latents, 5 tokens (2 context);
    with the published DiT-XL's MLP blocks and as the fork resolves it (spatial_mlp_ratio unset: no MLP, quirk Q2)."""
    cfg = bench.k600_cfg(sampling_timesteps=2, spatial_mlp_ratio=mlp)
    g = torch.Generator().manual_seed(123)
    xs = torch.randn((1, 5, 16, 16, 16), generator=g)
    _rollout(cfg, xs, None, 2, seed=43)


@pytest.mark.parametrize("latency_mode", [False, True])
def test_dmlab_fullsize_rollout_vs_oracle(latency_mode):
    """BASELINE configs[4] at batch 1: DiT3D-B (12 x 768, 12 heads of 64, patch 2, MLP x4) on 16 frames of 32x8x8 latents
    (256 token rows) with action conditioning — through the default kernels and in latency mode (ops.set_latency_mode:
    split-K GEMMs + the gated residual fused into the next AdaLN, block-per-row AdaLN, single-tile attention items)."""
    from dfot_b200 import ops
    ops.set_latency_mode(latency_mode)
    try:
        cfg = bench.dmlab_cfg(sampling_timesteps=2, frames=16)
        g = torch.Generator().manual_seed(5)
        xs = torch.randn((1, 16, 32, 8, 8), generator=g)
        conds = torch.randn((1, 16, 3), generator=g)
        _rollout(cfg, xs, conds, 4, seed=53)
    finally:
        ops.set_latency_mode(False)


def test_k600_fullsize_vanilla_hg_batch2():
    """Same backbone under vanilla history guidance (2 branch rows per sample) at batch 2: the batched-row path of the
    28-block network, not only B = 1."""
    cfg = bench.k600_cfg(sampling_timesteps=2)
    cfg["tasks"]["prediction"]["history_guidance"] = dict(name="vanilla", guidance_scale=2.0, visualize=False)
    g = torch.Generator().manual_seed(7)
    xs = torch.randn((2, 5, 16, 16, 16), generator=g)
    _rollout(cfg, xs, None, 2, seed=47)


def test_re10k_fullsize_forward_row_vs_oracle():
    """One forward-row of the full-size U-ViT3DPose with the pose masked (the unconditional guidance branch skips the
    pose cache) and one with it — the two row kinds of every vanilla-HG step — at a mid-range noise level."""
    from oracle.pose import ray_encoding
    from oracle.uvit3d_pose import UViT3DPoseOracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = bench.re10k_cfg(sampling_timesteps=2)
    algo = bench.make_weights(cfg, 0)
    model = algo.diffusion_model.model
    oracle = UViT3DPoseOracle(cfg["backbone"], cfg["x_shape"], 8, {k: v.detach().clone() for k, v in model.state_dict().items()})
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 8, 3, 256, 256), generator=g)
    levels = torch.randn((2, 8), generator=g) * 0.5
    poses = bench.synthetic_poses(1, 8)
    enc = ray_encoding(poses.repeat_interleave(2, 0), 256, "first", None, "ray_encoding")
    mask = torch.tensor([True, False])
    with torch.no_grad():
        ref = oracle(x, levels, enc, mask)
    del enc
    algo = algo.to(DEV).eval()
    cond = algo._window_conditions(poses.to(DEV), 2)
    out = algo.diffusion_model.model(x.to(DEV), levels.to(DEV), cond, mask.to(DEV)).cpu()
    err = (out - ref).abs().max().item()
    print(f"full-size U-ViT3DPose forward rows: max-abs {err:.3e} (|ref| max {ref.abs().max().item():.3f})")
    assert err <= STEP_TOL
