"""TEST INFRASTRUCTURE — the oracle against the EXECUTED reference at the two FULL-SIZE benchmarked configurations
(bench.re10k_cfg(): U-ViT3DPose 44 blocks, 8 frames 256x256, vanilla history guidance; bench.k600_cfg(): DiT3D-XL 28 x 1152):
a one-DDIM-step rollout through the reference's public `_predict_videos` and through oracle.sampler.SamplerOracle on the same
weights (the PRODUCT's state dict, loaded strictly into the reference: its keys are the reference's), inputs and torch seed.
Closes the chain GPU == oracle (tests/test_gpu_fullsize_parity.py) == reference at the sizes the numbers are quoted on.
Runs wherever the reference is present (/root/reference, or oracle/_ref/reference), in a process of its own:
    python -m oracle.check_fullsize [re10k] [k600] [dmlab]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
from oracle import ref_shim  # noqa: E402

TOL = 2e-4          # fp32 against fp32: summation-order differences over 28-44 residual blocks


def check(name: str) -> float:
    from helpers import build_oracle
    from algorithms.dfot.dfot_video import DFoTVideo as RefVideo
    from algorithms.dfot.dfot_video_pose import DFoTVideoPose as RefVideoPose
    g = torch.Generator().manual_seed(123)
    if name == "re10k":
        cfg = bench.re10k_cfg(sampling_timesteps=1)
        xs = torch.rand((1, 8, 3, 256, 256), generator=g)
        mean = torch.tensor(cfg["data_mean"]).reshape(1, 1, 3, 1, 1)
        xs = (xs - mean) / torch.tensor(cfg["data_std"]).reshape(1, 1, 3, 1, 1)
        conds, n_ctx = bench.synthetic_poses(1, 8), 1
    elif name == "dmlab":                                          # BASELINE configs[4]: DiT-B, action-conditioned, T = 16
        cfg = bench.dmlab_cfg(sampling_timesteps=1, frames=16)
        xs, conds, n_ctx = torch.randn((1, 16, 32, 8, 8), generator=g), torch.randn((1, 16, 3), generator=g), 4
    else:
        cfg = bench.k600_cfg(sampling_timesteps=1)
        xs, conds, n_ctx = torch.randn((1, 5, 16, 16, 16), generator=g), None, 2
    algo = bench.make_weights(cfg, 0)                              # the product class, random-init as in the bench
    sd = {k: v.detach().clone() for k, v in algo.state_dict().items()}
    del algo
    rcfg = json.loads(json.dumps(cfg))
    if rcfg["latent"]["enabled"] and rcfg["latent"]["downsampling_factor"][0] > 1:
        rcfg["latent"]["type"] = "online"                          # kinetics_600.yaml:9 (asserted by the reference)
    ref = (RefVideoPose if name == "re10k" else RefVideo)(ref_shim.to_dc(rcfg)).eval()
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("vae.", "metrics")) for k in missing), (missing[:5], unexpected[:5])
    torch.manual_seed(77)
    with torch.no_grad():
        want = ref._predict_videos(xs.clone(), n_context_tokens=n_ctx, conditions=conds)
    weights = {k[len("diffusion_model.model."):]: v for k, v in sd.items() if k.startswith("diffusion_model.model.")}
    oracle, _ = build_oracle(json.loads(json.dumps(cfg)), weights)
    torch.manual_seed(77)
    with torch.no_grad():
        got = oracle.predict_videos(xs.clone(), n_ctx, conds)
    err = (got - want).abs().max().item()
    print(f"{name}: |oracle - reference| = {err:.2e} over a 1-step rollout (output scale {want.abs().max().item():.2f})")
    assert torch.equal(got[:, :n_ctx], want[:, :n_ctx]) and err <= TOL * max(1.0, want.abs().max().item())
    return err


def main() -> int:
    torch.set_num_threads(os.cpu_count() or 1)
    ref_shim.install()
    for name in (sys.argv[1:] or ["k600", "dmlab", "re10k"]):
        check(name)
    print("OK")
    return 0


if __name__ == "__main__":
    sys.exit(main())
