"""TEST INFRASTRUCTURE — golden rollouts of the DiT3D absolute-position ablations that run in the fork
(`backbone.pos_emb_type: learned_1d | sinusoidal_1d`, dit_base.py:230-253, 352-353, 504-525) by EXECUTING the reference
(authoring container only):
    python -m oracle.make_goldens_posemb
Writes tests/golden/case_{learned_1d,sinusoidal_1d}_vanilla.{npz,json} and tests/golden/weights_{learned_1d,sinusoidal_1d}.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_goldens as mg  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.cases import _small  # noqa: E402


def posemb_cases():
    hg = {"tasks.prediction.history_guidance": dict(name="vanilla", guidance_scale=2.0, visualize=False)}
    return {f"{kind}_vanilla": dict(cfg=_small(**{**hg, "backbone.pos_emb_type": kind}), batch=2, weights=kind)
            for kind in ("learned_1d", "sinusoidal_1d")}


def main():
    ref_shim.install()
    weights = {}
    for name, spec in posemb_cases().items():
        mg.run_case(name, spec, weights)
    for w, sd in weights.items():
        np.savez_compressed(os.path.join(mg.OUT, f"weights_{w}.npz"), **sd)
        print(w, [k for k in sd if "pos_emb" in k])


if __name__ == "__main__":
    main()
