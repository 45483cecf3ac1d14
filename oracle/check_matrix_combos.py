"""TEST INFRASTRUCTURE — the oracle's DiT3D restatement against the EXECUTED reference backbone on the matrix-attention
combinations that no golden rollout covers (oracle/cases.py MATRIX_COMBOS: block type x head grouping x RoPE mode x bias).
Runs wherever the reference is present — /root/reference in the authoring container, or its unmodified copy under
oracle/_ref/reference (oracle/build_ref.py) — in a process of its own, because oracle/ref_shim.py installs stand-in modules:
    python -m oracle.check_matrix_combos
Prints one line per combination and exits non-zero on a mismatch above 1e-5."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.cases import MATRIX_COMBOS, matrix_combo_cfg  # noqa: E402
from oracle.dit3d import DiT3DOracle  # noqa: E402


def main() -> int:
    ref_shim.install()
    from algorithms.dfot.backbones.dit.dit3d import DiT3D        # the reference's class
    worst = 0.0
    for i, combo in enumerate(MATRIX_COMBOS):
        cfg = matrix_combo_cfg(combo)
        torch.manual_seed(100 + i)
        ref = DiT3D(cfg=ref_shim.to_dc(cfg), x_shape=[4, 8, 8], max_tokens=4, external_cond_type=None,
                    external_cond_num_classes=None, external_cond_dim=0, use_causal_mask=False).eval()
        ref_shim.rerandomize_zero_params(ref, 7)
        sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        oracle = DiT3DOracle(cfg, [4, 8, 8], 4, sd)
        g = torch.Generator().manual_seed(200 + i)
        x, lv = torch.randn((2, 4, 4, 8, 8), generator=g), torch.randint(0, 1000, (2, 4), generator=g)
        with torch.no_grad():
            want = ref(x, lv)
        err = (oracle(x, lv) - want).abs().max().item()
        worst = max(worst, err)
        print(f"combo {i} {combo}: |oracle - reference| = {err:.2e} (output scale {want.abs().max().item():.2f})")
        if not err <= 1e-5 * max(1.0, want.abs().max().item()):
            return 1
    print(f"OK: {len(MATRIX_COMBOS)} combinations, worst {worst:.2e}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
