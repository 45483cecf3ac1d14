"""-m gpu: the experiment / validation driver and checkpoint ingestion on the CUDA path:
a Lightning-style `.ckpt` (EMA list + torch.compile key prefixes) is written from the reference's golden weights, loaded
through `SamplingExperiment`, and the driver's rollout must reproduce the fixture of the executed reference.
Reference: experiments/simple_video_generation.py:324-487, 602-629; base_pytorch_video_algo.py:1096-1201."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from dfot_b200.experiments import SamplingExperiment  # noqa: E402
from helpers import build_product, load_case  # noqa: E402

DEV = torch.device("cuda", 0)


def _lightning_ckpt(path, cfg, weights):
    """state_dict with `_orig_mod.` prefixes and ZEROED tensors; the real weights only in the EMA list, in the order of the
    reference's named_parameters() (= the key order of the golden weight file, buffers skipped)."""
    algo = build_product(cfg)
    names = [k for k, _ in algo.diffusion_model.named_parameters()]
    sd = {("diffusion_model._orig_mod." + k[len("diffusion_model."):]): torch.zeros_like(v)
          for k, v in algo.state_dict().items() if k.startswith("diffusion_model.")}
    for k, v in weights.items():          # persistent buffers (Fourier frequencies) live in the state_dict proper
        if "model." + k not in names:
            sd["diffusion_model._orig_mod.model." + k] = v
    ema = [weights[n[len("model."):]] for n in names]
    torch.save({"state_dict": sd, "optimizer_states": [{"ema": ema}], "epoch": 7, "global_step": 1234}, path)


@pytest.mark.parametrize("case", ["uvit_pose_vanilla", "continuous_action"])
def test_driver_loads_a_lightning_ckpt_and_reproduces_the_reference_rollout(case, tmp_path):
    meta, arr, weights = load_case(case)
    cfg = meta["cfg"]
    ckpt = str(tmp_path / "model.ckpt")
    _lightning_ckpt(ckpt, cfg, weights)
    exp = SamplingExperiment(cfg, ckpt, DEV, manual_seed=None)
    algo = exp.algo
    torch.manual_seed(meta["sampling_seed"])
    algo.diffusion_model.noise_source = lambda shape, device: torch.randn(shape).to(device)
    xs = torch.from_numpy(arr["xs"])
    videos = algo._unnormalize_x(xs.to(DEV)).cpu()           # the driver takes dataset-space videos
    batch = {"videos": videos}
    if "conds" in arr:
        batch["conds"] = torch.from_numpy(arr["conds"])
    out = exp.run_validation([batch])[0]
    ref = algo._unnormalize_x(torch.from_numpy(arr["prediction"]).to(DEV)).cpu()
    n = cfg["context_frames"]
    pred = out["prediction"].cpu()
    rng = (ref.max() - ref.min()).item()
    mse = ((pred[:, n:] - ref[:, n:]) ** 2).mean().item()
    assert 10 * math.log10(rng * rng / max(mse, 1e-30)) >= 40.0
    assert exp.stats["forward_rows"] == int(arr["n_steps"]) * 2 * xs.shape[0] and exp.stats["videos"] == xs.shape[0]
    assert torch.allclose(out["gt"].cpu(), videos, atol=1e-6)


def test_safetensors_release_checkpoint_on_gpu(tmp_path):
    from safetensors.torch import save_file
    meta, arr, weights = load_case("vanilla")
    cfg = meta["cfg"]
    path = str(tmp_path / "release.safetensors")
    save_file({"diffusion_model.model." + k: v.contiguous() for k, v in weights.items()}, path)
    exp = SamplingExperiment(cfg, path, DEV)
    torch.manual_seed(meta["sampling_seed"])
    exp.algo.diffusion_model.noise_source = lambda shape, device: torch.randn(shape).to(device)
    out = exp.algo._predict_videos(torch.from_numpy(arr["xs"]).to(DEV), cfg["context_frames"], None).cpu()
    assert np.abs(out.numpy() - arr["prediction"]).max() <= 5e-2 and exp.algo.nfe_rows == int(arr["n_steps"]) * 4
