"""TEST INFRASTRUCTURE — golden rollouts of the DDPM ancestral sampler (`sampling_timesteps == timesteps`,
discrete_diffusion.py:403-452) by EXECUTING the reference (authoring container only):
    python -m oracle.make_goldens_ddpm
Writes tests/golden/case_ddpm_vanilla.{npz,json} (20-step schedule, vanilla history guidance: two branch rows with
their own step noise) and case_ddpm_pyramid.{npz,json} (autoregressive scheduling matrix: frames that wait at the top
level are NOT kept by the DDPM step — only level -1 is).  Both reuse the `plain` weights (same architecture and seeds)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_goldens as mg  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.cases import _small  # noqa: E402


def ddpm_cases():
    ddpm = {"diffusion.timesteps": 20, "diffusion.sampling_timesteps": 20}
    return {
        "ddpm_vanilla": dict(cfg=_small(**{**ddpm, "tasks.prediction.history_guidance":
                                           dict(name="vanilla", guidance_scale=2.0, visualize=False)}),
                             batch=2, weights="plain"),
        "ddpm_pyramid": dict(cfg=_small(**{**ddpm, "scheduling_matrix": "autoregressive"}), batch=1, weights="plain"),
    }


def main():
    ref_shim.install()
    plain = dict(np.load(os.path.join(mg.OUT, "weights_plain.npz")))
    for name, spec in ddpm_cases().items():
        weights = {"plain": plain}        # run_case asserts the freshly built reference weights equal the committed file
        mg.run_case(name, spec, weights)


if __name__ == "__main__":
    main()
