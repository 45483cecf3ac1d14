"""TEST INFRASTRUCTURE — goldens of the fork's refinement sampling (`refinement_sampling.enabled`, dfot_video.py:765-1008,
base_pytorch_video_algo.py:949-976) by EXECUTING the reference (authoring container only):
    python -m oracle.make_goldens_refine
Writes tests/golden/refine_matrices.json (integer walks) and tests/golden/case_refine_conditional.{npz,json} (a rollout
with per-step traces, continuous diffusion; weights in tests/golden/weights_plain_fourier.npz)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_goldens as mg  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.cases import _small, algorithm_cfg, continuous_overrides  # noqa: E402

MATRICES = [dict(steps=6, goback_length=2, n_goback=2, horizon=4, padding=0),
            dict(steps=10, goback_length=3, n_goback=1, horizon=3, padding=2),
            dict(steps=50, goback_length=20, n_goback=5, horizon=8, padding=0),
            dict(steps=5, goback_length=2, n_goback=1, horizon=2, padding=1),
            dict(steps=4, goback_length=4, n_goback=3, horizon=2, padding=0)]


def refine_case():
    # continuous (shifted-cosine) schedule: with the discrete cosine schedule alphas_cumprod[-1] == 0.0, so the re-noising
    # scale alphas_cumprod[-1] / alphas_cumprod[-1] of every context token (level -1) is 0/0 and the reference's
    # refinement rollout is NaN from the first go-back on (quirk Q11, probed) — not a case worth pinning
    return dict(cfg=_small(**{**continuous_overrides(), "diffusion.sampling_timesteps": 6,
                              "refinement_sampling": dict(enabled=True, goback_length=2, n_goback=2)}),
                batch=2, weights="plain_fourier")


def main():
    ref_shim.install()
    mats = []
    for m in MATRICES:
        cfg = algorithm_cfg(**{"backbone.hidden_size": 64, "backbone.depth": 1, "backbone.num_heads": 1, "x_shape": [4, 8, 8],
                               "diffusion.sampling_timesteps": m["steps"]})
        algo = mg.build_reference_algo(cfg)
        S = algo._generate_refine_scheduling_matrix(m["horizon"], m["goback_length"], m["n_goback"], m["padding"])
        mats.append(dict(**m, matrix=S.tolist()))
    with open(os.path.join(mg.OUT, "refine_matrices.json"), "w") as f:
        json.dump(mats, f)
    weights = {}
    mg.run_case("refine_conditional", refine_case(), weights)
    for w, sd in weights.items():       # the Fourier noise embedding has its own buffers: a weight set of its own
        np.savez_compressed(os.path.join(mg.OUT, f"weights_{w}.npz"), **sd)
    print("matrices:", [len(m["matrix"]) for m in mats])


if __name__ == "__main__":
    main()
