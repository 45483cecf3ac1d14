"""Shared test helpers: golden loading and oracle construction (test infrastructure)."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_case(name):
    with open(os.path.join(GOLDEN, f"case_{name}.json")) as f:
        meta = json.load(f)
    arrays = dict(np.load(os.path.join(GOLDEN, f"case_{name}.npz")))
    weights = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, f"weights_{meta['weights']}.npz")).items()}
    return meta, arrays, weights


def case_names():
    return sorted(f[len("case_"):-len(".json")] for f in os.listdir(GOLDEN) if f.startswith("case_") and f.endswith(".json"))


def build_oracle(cfg, weights, randn=torch.randn, randn_like=torch.randn_like):
    from oracle.dit3d import DiT3DOracle
    from oracle.sampler import SamplerOracle
    probe = SamplerOracle(cfg, None)
    if cfg["backbone"]["name"] == "u_vit3d_pose":
        from oracle.uvit3d_pose import UViT3DPoseOracle
        model = UViT3DPoseOracle(cfg["backbone"], probe.x_shape, probe.max_tokens, weights)
    else:
        model = DiT3DOracle(cfg["backbone"], probe.x_shape, probe.max_tokens, weights,
                            external_cond_dim=probe.external_cond_dim)
    return SamplerOracle(cfg, model, randn, randn_like), model


def build_product(cfg):
    """The product algorithm class for a golden case's config."""
    from dfot_b200.algorithms.dfot import DFoTVideo, DFoTVideoPose
    return (DFoTVideoPose if cfg["backbone"]["name"] == "u_vit3d_pose" else DFoTVideo)(cfg)


class NoiseBank:
    """Deterministic noise shared by oracle (CPU) and product (GPU): call i of shape s -> randn(seed+i)."""

    def __init__(self, seed=1234, device="cpu"):
        self.seed, self.i, self.device = seed, 0, device

    def randn(self, shape, **kw):
        g = torch.Generator().manual_seed(self.seed + self.i)
        self.i += 1
        return torch.randn(tuple(shape), generator=g).to(self.device)

    def randn_like(self, x):
        return self.randn(x.shape)

    # usable directly as `diffusion_model.noise_source` — and forkable, which the lockstep rounds of the multi-GPU path need
    def __call__(self, shape, device):
        return self.randn(shape).to(device)

    def get_state(self):
        return self.i

    def set_state(self, i):
        self.i = i


from oracle.cases import MATRIX_COMBOS, matrix_combo_cfg  # noqa: E402,F401


def matrix_combo_model(combo, seed=0):
    """(product DiT3D with redrawn zero-initialised parameters, oracle on the same weights, x, levels) for one combo."""
    from oracle.dit3d import DiT3DOracle
    from dfot_b200.algorithms.dfot.backbones.dit.dit3d import DiT3D
    cfg = matrix_combo_cfg(combo)
    torch.manual_seed(seed)
    model = DiT3D(cfg, [4, 8, 8], 4, use_causal_mask=False).eval()
    for prm in model.parameters():                       # AdaLN-Zero, final layer and the matrix biases start at zero
        if prm.abs().sum() == 0:
            torch.nn.init.normal_(prm, std=0.05)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    oracle = DiT3DOracle(cfg, [4, 8, 8], 4, sd)
    g = torch.Generator().manual_seed(seed + 1)
    return model, oracle, torch.randn((2, 4, 4, 8, 8), generator=g), torch.randint(0, 1000, (2, 4), generator=g)
