"""TEST INFRASTRUCTURE — golden rollouts of the matrix-attention DiT3D variants (`algorithm/backbone=dit3d_factorized_matrix`
and `dit3d_full_matrix`: variant factorized_matrix_attention / full_matrix_attention, matrix_block=matrix, sinusoidal_2d
positions; dit_blocks.py:211-350, 549-652, dit_base.py:129-222, 396-405) by EXECUTING the reference (authoring container
only):
    python -m oracle.make_goldens_matrix
Writes tests/golden/case_matrix_*.{npz,json} and tests/golden/weights_matrix_*.npz, and checks that with one row per column
head (embed_col_dim == num_col_heads, every shipped configuration) `flatten_matrix_rope` and `matrix_multi_token` do not
change the computation (bit-identical rollouts) — the product accepts them on that ground."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_goldens as mg  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.cases import _small  # noqa: E402


def matrix_case(variant: str, weights: str, batch: int, hg=None, **kw):
    """The shipped dit3d_factorized_matrix.yaml at golden size: E = 64, one row head of 64, one column of one row."""
    o = {"backbone.variant": variant, "backbone.pos_emb_type": "sinusoidal_2d", "backbone.use_temporal_rope": True,
         "backbone.hidden_size": None, "backbone.embed_col_dim": 1, "backbone.embed_row_dim": 64,
         "backbone.num_col_heads": 1, "backbone.num_row_heads": 1, "backbone.mlp_ratio": 2.0,
         "backbone.spatial_mlp_ratio": 2.0, "backbone.use_bias": False, "backbone.matrix_block": "matrix",
         "backbone.flatten_matrix_rope": False, "backbone.matrix_multi_token": False, "backbone.fixed_u": None}
    if hg is not None:
        o["tasks.prediction.history_guidance"] = hg
    o.update(kw)
    return dict(cfg=_small(**o), batch=batch, weights=weights)


def cases():
    vanilla = dict(name="vanilla", guidance_scale=2.0, visualize=False)
    return {
        "matrix_factorized_vanilla": matrix_case("factorized_matrix_attention", "matrix_factorized", 2, vanilla),
        "matrix_full_cols2": matrix_case("full_matrix_attention", "matrix_full_cols2", 1, None,
                                         **{"backbone.embed_col_dim": 2, "backbone.num_col_heads": 2,
                                            "backbone.hidden_size": 64}),
        "matrix_factorized_bias": matrix_case("factorized_matrix_attention", "matrix_factorized_bias", 1, None,
                                              **{"backbone.use_bias": True, "backbone.use_temporal_rope": False}),
        # the other two entries of dit_base.py:27-31 `matrix_blocks`
        "matrix_self_factorized": matrix_case("factorized_matrix_attention", "matrix_self", 1, None,
                                              **{"backbone.matrix_block": "matrix_self"}),
        "matrix_cross_full": matrix_case("full_matrix_attention", "matrix_cross", 1, vanilla,
                                         **{"backbone.matrix_block": "matrix_cross", "backbone.hidden_size": 64,
                                            "backbone.embed_col_dim": 2, "backbone.num_col_heads": 2}),
        # more than one row per column head (dit_blocks.py:312-340): every row its own sequence (multi-token), the flattened
        # [n, d] feature rotated as one vector, and — with two row heads of 32 — rotated row by row
        "matrix_rows2_multi_token": matrix_case("full_matrix_attention", "matrix_rows4", 1, None,
                                                **{"backbone.embed_col_dim": 4, "backbone.num_col_heads": 2,
                                                   "backbone.hidden_size": 64, "backbone.matrix_multi_token": True}),
        "matrix_rows2_flatten": matrix_case("full_matrix_attention", "matrix_rows4", 1, None,
                                            **{"backbone.embed_col_dim": 4, "backbone.num_col_heads": 2,
                                               "backbone.hidden_size": 64, "backbone.flatten_matrix_rope": True}),
        # one bias row per column row (qkv_bias [embed_col_dim, 3E], dit_blocks.py:283-286, 303-304)
        "matrix_bias_cols2": matrix_case("full_matrix_attention", "matrix_bias_cols2", 1, None,
                                         **{"backbone.embed_col_dim": 2, "backbone.num_col_heads": 2,
                                            "backbone.hidden_size": 64, "backbone.use_bias": True}),
        "matrix_rows2_grouped": matrix_case("factorized_matrix_attention", "matrix_rows2_grouped", 1, vanilla,
                                            **{"backbone.embed_col_dim": 2, "backbone.num_col_heads": 1,
                                               "backbone.num_row_heads": 2}),
    }


def rollout(spec):
    algo = mg.build_reference_algo(spec["cfg"])
    xs, conds = mg.synthetic_inputs(spec["cfg"], spec["batch"])
    torch.manual_seed(mg.SAMPLING_SEED)
    with torch.no_grad():
        return algo._predict_videos(xs.clone(), n_context_tokens=spec["cfg"]["context_frames"], conditions=conds).numpy()


def main():
    ref_shim.install()
    weights = {}
    for name, spec in cases().items():
        mg.run_case(name, spec, weights)
    for w, sd in weights.items():
        np.savez_compressed(os.path.join(mg.OUT, f"weights_{w}.npz"), **sd)
        print(w, len(sd), "tensors;", sorted(k for k in sd if ".attn.qkv_" in k or ".attn.proj_" in k)[:6])
    ref = np.load(os.path.join(mg.OUT, "case_matrix_full_cols2.npz"))["prediction"]
    for flag in ("backbone.flatten_matrix_rope", "backbone.matrix_multi_token"):
        spec = matrix_case("full_matrix_attention", "matrix_full_cols2", 1, None,
                           **{"backbone.embed_col_dim": 2, "backbone.num_col_heads": 2, "backbone.hidden_size": 64,
                              flag: True})
        assert np.array_equal(rollout(spec), ref), flag
        print(f"{flag}=True == the plain rollout with one row per column head (bit-identical)")


if __name__ == "__main__":
    main()
