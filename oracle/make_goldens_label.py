"""TEST INFRASTRUCTURE — golden rollout with class-label conditioning (`external_cond_type: label`, the cond_ucf_101
configurations; base_backbone.py:47-51, dit3d.py:171-173) by EXECUTING the reference (authoring container only):
    python -m oracle.make_goldens_label
Writes tests/golden/case_label_vanilla.{npz,json} and tests/golden/weights_label.npz.  Labels are one int64 per clip,
shape [B, 1], as the reference's dataset yields them (datasets/video/ucf_101.py:304-309)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_goldens as mg  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.cases import _small  # noqa: E402

NUM_CLASSES = 10


def label_case():
    return dict(cfg=_small(**{"external_cond_type": "label", "external_cond_num_classes": NUM_CLASSES, "external_cond_dim": 1,
                              "backbone.external_cond_dropout": 0.1,
                              "tasks.prediction.history_guidance": dict(name="vanilla", guidance_scale=2.0, visualize=False)}),
                batch=2, weights="label")


def label_inputs(cfg: dict, batch: int):
    g = torch.Generator().manual_seed(mg.DATA_SEED)
    xs = torch.randn((batch, cfg["n_frames"], *cfg["x_shape"]), generator=g)
    return xs, torch.randint(0, NUM_CLASSES, (batch, 1), generator=g)


def main():
    ref_shim.install()
    mg.synthetic_inputs = label_inputs
    weights = {}
    mg.run_case("label_vanilla", label_case(), weights)
    for w, sd in weights.items():
        np.savez_compressed(os.path.join(mg.OUT, f"weights_{w}.npz"), **sd)
        print(w, sorted(k for k in sd if "external_cond" in k), {k: v.shape for k, v in sd.items() if "external_cond" in k})


if __name__ == "__main__":
    main()
