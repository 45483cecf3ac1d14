"""Sampling driver: config → algorithm → checkpoint → batches → videos, on CPU with the kernel
contract emulations (the CUDA path is covered by the -m gpu tests)."""
import numpy as np
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
