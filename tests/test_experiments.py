"""Sampling driver: config → algorithm → checkpoint → batches → videos, on CPU with the kernel
contract emulations (the CUDA path is covered by the -m gpu tests)."""
import numpy as np
import pytest
import torch

import k4_emulation
import ops_emulation
from dfot_b200 import ops
from dfot_b200.experiments import SamplingExperiment
from helpers import load_case


def test_driver_matches_reference_rollout(monkeypatch, tmp_path):
    meta, arr, weights = load_case("uvit_pose_vanilla")
    cfg = meta["cfg"]
    ckpt = str(tmp_path / "m.ckpt")
    torch.save({"state_dict": {"diffusion_model.model." + k: v for k, v in weights.items()}, "pretrained_ema": True,
                "optimizer_states": []}, ckpt)
    ops_emulation.install(monkeypatch)
    monkeypatch.setattr(ops, "sampler_step_hg", k4_emulation.emulate)
    exp = SamplingExperiment(cfg, ckpt)
    exp.algo.model_in_dtype = torch.float32
    torch.manual_seed(meta["sampling_seed"])
    exp.algo.diffusion_model.noise_source = lambda shape, device: torch.randn(shape)
    # the driver takes dataset-space videos; the golden inputs are already normalised → un-normalise them first
    videos = exp.algo._unnormalize_x(torch.from_numpy(arr["xs"]))
    out = exp.run_validation([{"videos": videos, "conds": torch.from_numpy(arr["conds"])}])[0]
    assert set(out) == {"gt", "prediction"} and exp.stats["forward_rows"] == int(arr["n_steps"]) * 2
    ref = exp.algo._unnormalize_x(torch.from_numpy(arr["prediction"]))
    n_ctx = cfg["context_frames"]
    mse = float(((out["prediction"][:, n_ctx:] - ref[:, n_ctx:]) ** 2).mean())
    rng = float(ref.max() - ref.min())
    assert 10 * np.log10(rng * rng / max(mse, 1e-30)) >= 40.0
    assert torch.allclose(out["gt"], videos, atol=1e-6)


def test_driver_decodes_latents_with_the_configured_vae(monkeypatch, tmp_path):
    """Latent-video configuration (temporal downsampling 4): the driver loads the VideoVAE named by `vae.pretrained_path`
    with the reference's checkpoint rules and returns decoded frames (dfot_video.py:104-111), `gt` from the dataset's
    videos when present.  The sampler is stubbed to the identity — the decode plumbing is what is under test."""
    from oracle.cases import algorithm_cfg
    from oracle.video_vae import VideoVAEDecoderOracle, decoder_param_shapes, seeded_weights
    mult = (1, 2, 2, 2)
    sd = seeded_weights(decoder_param_shapes(32, 4, 4, mult), 11)
    vae_ckpt = str(tmp_path / "vae.ckpt")
    torch.save({"model_cfg": dict(hidden_size=32, z_channels=4, embed_dim=4, hidden_size_mult=list(mult), resolution=32,
                                  temporal_length=9),
                "optimizer_states": [], "state_dict": {f"vae.{k}": v for k, v in sd.items()}}, vae_ckpt)
    cfg = algorithm_cfg(**{"backbone.hidden_size": 64, "backbone.depth": 1, "backbone.num_heads": 1, "x_shape": [3, 32, 32],
                           "latent.enabled": True, "latent.downsampling_factor": [4, 8], "latent.num_channels": 4,
                           "max_frames": 9, "n_frames": 9, "context_frames": 5, "vae.pretrained_path": vae_ckpt,
                           "vae.batch_size": 2, "data_mean": [[[0.2]]] * 4, "data_std": [[[2.0]]] * 4})
    ops_emulation.install(monkeypatch)
    exp = SamplingExperiment(cfg, None)
    monkeypatch.setattr(exp.algo, "sample_sharded", lambda xs, conds, n_ctx: xs)
    latents = torch.randn((3, 3, 4, 4, 4), generator=torch.Generator().manual_seed(0))
    out = exp.run_validation([{"latents": latents}])[0]
    assert out["prediction"].shape == out["gt"].shape == (3, 9, 3, 32, 32)
    ref = VideoVAEDecoderOracle(sd, mult).decode(latents.permute(0, 2, 1, 3, 4), 9).permute(0, 2, 1, 3, 4) * 0.5 + 0.5
    for k in ("gt", "prediction"):
        assert ((out[k] - ref).norm() / ref.norm()).item() <= 2e-2
    gt = torch.rand((3, 9, 3, 32, 32))
    out = exp.run_validation([{"latents": latents, "videos": gt}])[0]
    assert torch.equal(out["gt"], gt)


def test_cli_resolves_the_reference_command_line(tmp_path):
    """`--config-dir <configurations> -- <python -m main arguments>`: composed like the reference's Hydra entry point."""
    import argparse
    import os
    from dfot_b200.experiments import resolve_cli_config
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "algorithm"))
    for rel, text in {"config.yaml": "defaults:\n  - experiment: gen\n  - dataset: d\n  - algorithm: a\nload: null\n",
                      "experiment/gen.yaml": "tasks: [training]\n", "dataset/d.yaml": "n_frames: 8\n",
                      "algorithm/a.yaml": "n_frames: ${dataset.n_frames}\nname: toy\n"}.items():
        os.makedirs(os.path.dirname(os.path.join(root, rel)), exist_ok=True)
        with open(os.path.join(root, rel), "w") as f:
            f.write(text)
    ns = argparse.Namespace(config=None, config_dir=root, ckpt=None,
                            overrides=["experiment.tasks=[validation]", "dataset.n_frames=200", "load=/w/model.ckpt"])
    tree, ckpt = resolve_cli_config(ns)
    assert tree == {"n_frames": 200, "name": "toy", "_name": "a"} and ckpt == "/w/model.ckpt"
    ns.overrides = ["load=pretrained:DFoT_RE10K.ckpt", "experiment.tasks=[validation]"]
    with pytest.raises(SystemExit, match="no network"):
        resolve_cli_config(ns)
    ns.overrides = []
    with pytest.raises(SystemExit, match="sampling tasks"):
        resolve_cli_config(ns)
