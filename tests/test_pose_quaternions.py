"""Quaternion arithmetic of the pose path (normalize_by="mean", `temporal` guidance pose fill-in).

roma (the reference's dependency, utils/geometry_utils.py:141-143, 172-205) is not installed; both restatements of its
three functions — oracle/roma_restatement.py (what the reference ran on when the fixtures were generated) and the
product's algorithms/dfot/pose_math.py — are pinned here against scipy.spatial.transform, and the product's camera table
against the CameraPose internals the reference itself produced (tests/golden/pose_quaternions.npz)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfot_b200.algorithms.dfot import pose_math  # noqa: E402
from dfot_b200.algorithms.dfot.dfot_video_pose import camera_table  # noqa: E402
from oracle import roma_restatement as rr  # noqa: E402

scipy_rot = pytest.importorskip("scipy.spatial.transform")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pose_quaternions.npz")


def _rotations(n, seed):
    rs = scipy_rot.Rotation.random(n, random_state=seed)
    # include half-turns and near-identity rotations: the largest-diagonal branches and the Taylor branches
    extra = scipy_rot.Rotation.from_rotvec(np.array([[np.pi, 0, 0], [0, np.pi - 1e-7, 0], [0, 0, np.pi], [1e-5, 0, 0],
                                                     [0, 2e-4, 1e-4], [0, 0, 0]]))
    return scipy_rot.Rotation.concatenate([rs, extra])


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_matrix_quaternion_mappings_match_scipy(impl):
    rs = _rotations(200, 3)
    M = torch.from_numpy(rs.as_matrix())
    to_q = rr.rotmat_to_unitquat if impl == "oracle" else pose_math.rotmat_to_quat
    to_m = rr.unitquat_to_rotmat if impl == "oracle" else pose_math.quat_to_rotmat
    q, qs = to_q(M).numpy(), rs.as_quat()
    assert np.minimum(np.abs(q - qs).max(1), np.abs(q + qs).max(1)).max() < 1e-12      # (q and -q are the same rotation)
    assert np.abs(to_m(torch.from_numpy(qs)).numpy() - rs.as_matrix()).max() < 1e-12
    # not normalising is part of the contract: a scaled quaternion gives the rotation scaled by |q|^2
    assert np.abs(to_m(torch.from_numpy(0.5 * qs)).numpy() - 0.25 * rs.as_matrix()).max() < 1e-12
    # batched shapes
    assert to_q(M.reshape(2, 103, 3, 3)).shape == (2, 103, 4)


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_slerp_matches_scipy(impl):
    rs = _rotations(64, 5)
    for i in range(0, 35):
        a, b = rs[i], rs[i + 35]
        t = np.linspace(0, 1, 6)
        ref = scipy_rot.Slerp([0, 1], scipy_rot.Rotation.concatenate([a, b]))(t).as_matrix()
        qa, qb, ts = torch.from_numpy(a.as_quat()), torch.from_numpy(b.as_quat()), torch.from_numpy(t)
        if impl == "oracle":
            out = rr.unitquat_to_rotmat(rr.unitquat_slerp(qa, qb, ts))
        else:
            out = pose_math.quat_to_rotmat(pose_math.quat_slerp(qa, qb, ts))
        assert np.abs(out.numpy() - ref).max() < 1e-9, i


def _table_parts(tab):
    return tab[..., 4:13].reshape(*tab.shape[:2], 3, 3), tab[..., 13:]


def test_camera_table_mean_matches_reference_camera_pose():
    g = np.load(GOLDEN)
    conds = torch.from_numpy(g["conds"])
    R_inv, origin = _table_parts(camera_table(conds, 8, "mean", None))
    R, T = torch.from_numpy(g["mean.R"]), torch.from_numpy(g["mean.T"])
    assert (R_inv - R.transpose(-1, -2)).abs().max() < 2e-6
    assert (origin + torch.einsum("btji,btj->bti", R, T)).abs().max() < 2e-5
    _, origin_b = _table_parts(camera_table(conds, 8, "mean", 1.0))
    Tb = torch.from_numpy(g["mean_bound.T"])
    assert (origin_b + torch.einsum("btji,btj->bti", R, Tb)).abs().max() < 2e-5
    assert float(Tb.abs().max()) == pytest.approx(1.0, abs=1e-6)


def test_camera_table_interpolated_poses_match_reference_camera_pose():
    g = np.load(GOLDEN)
    conds, mask = torch.from_numpy(g["conds"]), torch.from_numpy(g["interp.mask"])
    R, t = pose_math.interpolate_masked_poses(conds[..., 4:].reshape(3, 7, 3, 4)[..., :3].float(),
                                              conds[..., 4:].reshape(3, 7, 3, 4)[..., 3].float(), mask)
    assert (R - torch.from_numpy(g["interp.R"])).abs().max() < 2e-6
    assert (t - torch.from_numpy(g["interp.T"])).abs().max() < 2e-6
    # the unmasked row is untouched up to the quaternion round trip; masked frames differ from the input poses
    assert (R[2] - conds[2, :, 4:].reshape(7, 3, 4)[..., :3]).abs().max() < 2e-6
    assert (R[0, 1] - conds[0, 1, 4:].reshape(3, 4)[:, :3]).abs().max() > 1e-3
    R_inv, origin = _table_parts(camera_table(conds, 8, "first", None, interp_mask=mask))
    Rg, Tg = torch.from_numpy(g["interp_first.R"]), torch.from_numpy(g["interp_first.T"])
    assert (R_inv - Rg.transpose(-1, -2)).abs().max() < 2e-6
    assert (origin + torch.einsum("btji,btj->bti", Rg, Tg)).abs().max() < 2e-5


def test_oracle_ray_encoding_with_interpolation_matches_reference():
    from oracle.pose import ray_encoding
    g = np.load(GOLDEN)
    enc = ray_encoding(torch.from_numpy(g["conds"]), 8, "first", None, "ray_encoding",
                       interp_mask=torch.from_numpy(g["interp.mask"]))
    ref = torch.from_numpy(g["interp_first.encoding"]).permute(0, 1, 4, 2, 3)
    # sin(x * 2^14 * pi) amplifies fp32 rounding of x by ~5e4: compare where the argument is small, bound the rest
    assert (enc - ref).abs().max() < 0.2
    assert (enc - ref)[:, :, :6].abs().max() < 1e-3
