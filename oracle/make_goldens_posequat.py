"""TEST INFRASTRUCTURE — golden fixtures of the two pose options that go through roma's quaternion routines, produced by
EXECUTING the reference (authoring container only):
    python -m oracle.make_goldens_posequat
  camera_pose_conditioning.normalize_by = "mean"     (utils/geometry_utils.py:137-155)
  `temporal` history guidance with camera poses       (algorithms/dfot/dfot_video_pose.py:73-81 ->
                                                       geometry_utils.py:170-206 replace_with_interpolation)
roma itself is not installed; the reference runs on oracle/roma_restatement.py (roma's three functions restated, every
one cross-checked against scipy.spatial.transform below and again in tests/test_pose_quaternions.py).  Writes
  tests/golden/case_uvit_pose_mean_vanilla.{npz,json}, tests/golden/case_uvit_pose_temporal.{npz,json}
  tests/golden/pose_quaternions.npz   CameraPose internals (R, T) after the reference's normalize_by_mean and
                                      replace_with_interpolation on synthetic trajectories + quaternion known answers
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_goldens as mg  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle import roma_restatement as rr  # noqa: E402
from oracle.cases import algorithm_cfg, golden_cases, synthetic_poses  # noqa: E402


def pose_cases():
    base = golden_cases()["uvit_pose_vanilla"]["cfg"]
    import copy
    mean = copy.deepcopy(base)
    mean["camera_pose_conditioning"]["normalize_by"] = "mean"
    mean["camera_pose_conditioning"]["bound"] = 1.0
    temporal = copy.deepcopy(base)
    temporal["context_frames"] = 2
    temporal["max_frames"] = temporal["n_frames"] = 5
    temporal["tasks"]["prediction"]["history_guidance"] = dict(
        name="temporal", hist_subsequences=[[0], [1]], hist_weights=[1.5, 1.5], visualize=False)
    return {"uvit_pose_mean_vanilla": dict(cfg=mean, batch=1, weights="uvit_pose", algo="dfot_video_pose"),
            "uvit_pose_temporal": dict(cfg=temporal, batch=1, weights="uvit_pose", algo="dfot_video_pose")}


def scipy_crosscheck():
    from scipy.spatial.transform import Rotation, Slerp
    rs = Rotation.random(64, random_state=7)
    M = torch.from_numpy(rs.as_matrix())
    q = rr.rotmat_to_unitquat(M).numpy()
    qs = rs.as_quat()
    assert np.minimum(np.abs(q - qs).max(1), np.abs(q + qs).max(1)).max() < 1e-12
    assert np.abs(rr.unitquat_to_rotmat(torch.from_numpy(qs)).numpy() - rs.as_matrix()).max() < 1e-12
    for i in range(16):
        t = np.linspace(0, 1, 6)
        ref = Slerp([0, 1], Rotation.concatenate([rs[i], rs[i + 32]]))(t).as_matrix()
        out = rr.unitquat_slerp(torch.from_numpy(rs[i].as_quat()), torch.from_numpy(rs[i + 32].as_quat()),
                                torch.from_numpy(t))
        assert np.abs(rr.unitquat_to_rotmat(out).numpy() - ref).max() < 1e-12
    print("roma restatement == scipy on 64 random rotations / 16 slerps")


def camera_pose_internals():
    """R, T of the reference's CameraPose after its own normalisations (the product's camera table is checked against them)."""
    from utils.geometry_utils import CameraPose
    conds = synthetic_poses(3, 7)
    # make the trajectories less tame: add a roll so that the largest-diagonal branch of the quaternion mapping is hit
    out = {"conds": conds.numpy()}
    cp = CameraPose.from_vectors(conds.clone())
    cp.normalize_by_mean()
    out["mean.R"], out["mean.T"] = cp._R.numpy(), cp._T.numpy()
    cp.scale_within_bounds(1.0)
    out["mean_bound.T"] = cp._T.numpy()
    masks = torch.tensor([[0, 1, 1, 0, 1, 0, 0], [1, 1, 0, 0, 0, 1, 1], [0, 0, 0, 0, 0, 0, 0]], dtype=torch.bool)
    cp = CameraPose.from_vectors(conds.clone())
    cp.replace_with_interpolation(masks)
    out["interp.mask"], out["interp.R"], out["interp.T"] = masks.numpy(), cp._R.numpy(), cp._T.numpy()
    cp.normalize_by_first()
    out["interp_first.R"], out["interp_first.T"] = cp._R.numpy(), cp._T.numpy()
    rays = cp.rays(resolution=8).to_pos_encoding()[0]
    out["interp_first.encoding"] = rays.numpy()
    np.savez_compressed(os.path.join(mg.OUT, "pose_quaternions.npz"), **out)
    print("pose_quaternions.npz:", {k: v.shape for k, v in out.items()})


def main():
    scipy_crosscheck()
    ref_shim.install()
    camera_pose_internals()
    weights = {}
    for name, spec in pose_cases().items():
        mg.run_case(name, spec, weights)
    for w, sd in weights.items():
        path = os.path.join(mg.OUT, f"weights_{w}.npz")
        if os.path.exists(path):   # same architecture + seeds as the committed fixture: must be identical
            old = np.load(path)
            assert all(np.array_equal(old[k], v) for k, v in sd.items()), w
        else:
            np.savez_compressed(path, **sd)


if __name__ == "__main__":
    main()
