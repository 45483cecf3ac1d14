"""ORACLE (test infrastructure): restatement of the four `roma` functions the reference's pose path calls.

The reference pins roma==1.5.2.1 (requirements.txt) and uses it in utils/geometry_utils.py:141-143 (quaternion mean of
`normalize_by="mean"`) and :172-205 (`replace_with_interpolation`, the `temporal` history-guidance pose fill-in):
    roma.rotmat_to_unitquat, roma.unitquat_to_rotmat, roma.unitquat_slerp
roma is not installed here and is not vendored under /root/reference, so the real package cannot be executed.  roma's
mappings are ports of SciPy's `Rotation` (XYZW convention; its sources cite scipy/spatial/transform/rotation.py), and SciPy
IS installed: `tests/test_pose_quaternions.py` pins every function below against scipy.spatial.transform
(`Rotation.from_matrix(...).as_quat()`, `Rotation.from_quat(...).as_matrix()`, `Slerp`) on random rotations.  What scipy
cannot pin is one behaviour the reference relies on: `unitquat_to_rotmat` does NOT normalise its argument, so the
arithmetic mean of unit quaternions (norm < 1) maps to a rotation matrix scaled by |q|^2 (geometry_utils.py:142-143 keeps
that scale) — restated from roma's formula, which is the unit-quaternion formula of SciPy's `as_matrix` without its
normalisation; stated as "parity pinned to scipy + published formula" in DESIGN.md.

`install(module)` fills the `roma` stub of oracle/ref_shim.py with these functions so that the reference's own
CameraPose.normalize_by_mean / replace_with_interpolation can be executed for the golden fixtures.
"""
import torch


def _flatten(t: torch.Tensor, end_dim: int):
    batch_shape = t.shape[: end_dim + 1] if end_dim >= 0 else t.shape[: t.dim() + end_dim + 1]
    return t.reshape((-1,) + tuple(t.shape[len(batch_shape):])), batch_shape


def rotmat_to_unitquat(R: torch.Tensor) -> torch.Tensor:
    """(..., 3, 3) rotation matrices -> (..., 4) unit quaternions, XYZW (roma.mappings.rotmat_to_unitquat = SciPy's
    `Rotation.from_matrix`: pick the largest of (R00, R11, R22, trace), build the quaternion from it, normalise)."""
    m, batch = _flatten(R, -3)
    n = m.shape[0]
    dec = torch.empty((n, 4), dtype=m.dtype, device=m.device)
    dec[:, :3] = m.diagonal(dim1=1, dim2=2)
    dec[:, -1] = dec[:, :3].sum(dim=1)
    choice = dec.argmax(dim=1)
    q = torch.empty((n, 4), dtype=m.dtype, device=m.device)
    ind = torch.nonzero(choice != 3, as_tuple=True)[0]
    i = choice[ind]
    j = (i + 1) % 3
    k = (j + 1) % 3
    q[ind, i] = 1 - dec[ind, -1] + 2 * m[ind, i, i]
    q[ind, j] = m[ind, j, i] + m[ind, i, j]
    q[ind, k] = m[ind, k, i] + m[ind, i, k]
    q[ind, 3] = m[ind, k, j] - m[ind, j, k]
    ind = torch.nonzero(choice == 3, as_tuple=True)[0]
    q[ind, 0] = m[ind, 2, 1] - m[ind, 1, 2]
    q[ind, 1] = m[ind, 0, 2] - m[ind, 2, 0]
    q[ind, 2] = m[ind, 1, 0] - m[ind, 0, 1]
    q[ind, 3] = 1 + dec[ind, -1]
    q = q / torch.norm(q, dim=1)[:, None]
    return q.reshape(tuple(batch) + (4,))


def unitquat_to_rotmat(quat: torch.Tensor) -> torch.Tensor:
    """(..., 4) XYZW -> (..., 3, 3); the unit-quaternion formula, NOT normalised (see the module docstring)."""
    x, y, z, w = quat[..., 0], quat[..., 1], quat[..., 2], quat[..., 3]
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    m = torch.empty(quat.shape[:-1] + (3, 3), dtype=quat.dtype, device=quat.device)
    m[..., 0, 0] = x2 - y2 - z2 + w2
    m[..., 1, 0] = 2 * (xy + zw)
    m[..., 2, 0] = 2 * (xz - yw)
    m[..., 0, 1] = 2 * (xy - zw)
    m[..., 1, 1] = -x2 + y2 - z2 + w2
    m[..., 2, 1] = 2 * (yz + xw)
    m[..., 0, 2] = 2 * (xz + yw)
    m[..., 1, 2] = 2 * (yz - xw)
    m[..., 2, 2] = -x2 - y2 + z2 + w2
    return m


def quat_conjugation(q: torch.Tensor) -> torch.Tensor:
    return torch.cat((-q[..., :3], q[..., 3:]), dim=-1)


def quat_product(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """Hamilton product, XYZW."""
    vec = p[..., None, 3] * q[..., :3] + q[..., None, 3] * p[..., :3] + torch.cross(p[..., :3], q[..., :3], dim=-1)
    last = p[..., 3] * q[..., 3] - torch.sum(p[..., :3] * q[..., :3], dim=-1)
    return torch.cat((vec, last[..., None]), dim=-1)


def unitquat_to_rotvec(quat: torch.Tensor, shortest_arc: bool = True) -> torch.Tensor:
    """SciPy's `as_rotvec`: angle = 2 atan2(|v|, w) with w >= 0 enforced, Taylor scale below 1e-3 rad."""
    q, batch = _flatten(quat, -2)
    q = q.clone()
    if shortest_arc:
        q[q[:, 3] < 0] *= -1
    half = torch.atan2(torch.norm(q[:, :3], dim=1), q[:, 3])
    angle = 2 * half
    small = torch.abs(angle) <= 1e-3
    scale = torch.empty(q.shape[0], dtype=q.dtype, device=q.device)
    scale[small] = 2 + angle[small] ** 2 / 12 + 7 * angle[small] ** 4 / 2880
    scale[~small] = angle[~small] / torch.sin(half[~small])
    return (scale[:, None] * q[:, :3]).reshape(tuple(batch) + (3,))


def rotvec_to_unitquat(rotvec: torch.Tensor) -> torch.Tensor:
    """SciPy's `from_rotvec`."""
    r, batch = _flatten(rotvec, -2)
    norms = torch.norm(r, dim=1)
    small = norms <= 1e-3
    scale = torch.empty(r.shape[0], dtype=r.dtype, device=r.device)
    scale[small] = 0.5 - norms[small] ** 2 / 48 + norms[small] ** 4 / 3840
    scale[~small] = torch.sin(norms[~small] / 2) / norms[~small]
    q = torch.empty((r.shape[0], 4), dtype=r.dtype, device=r.device)
    q[:, :3] = scale[:, None] * r
    q[:, 3] = torch.cos(norms / 2)
    return q.reshape(tuple(batch) + (4,))


def unitquat_slerp(q0: torch.Tensor, q1: torch.Tensor, steps: torch.Tensor, shortest_arc: bool = True) -> torch.Tensor:
    """q(s) = q0 * exp(s * log(q0^-1 q1)) for every s in `steps`; result (steps..., batch..., 4)."""
    rel_rotvec = unitquat_to_rotvec(quat_product(quat_conjugation(q0), q1), shortest_arc=shortest_arc)
    rel = steps.reshape(steps.shape + (1,) * rel_rotvec.dim()) * rel_rotvec.reshape((1,) * steps.dim() + rel_rotvec.shape)
    rots = rotvec_to_unitquat(rel.reshape(-1, 3)).reshape(*rel.shape[:-1], 4)
    base = q0.reshape((1,) * steps.dim() + q0.shape).repeat(steps.shape + (1,) * q0.dim())
    return quat_product(base, rots)


def install(module) -> None:
    """Fill a stub module named `roma` (oracle/ref_shim.py) with the restated functions."""
    for fn in (rotmat_to_unitquat, unitquat_to_rotmat, unitquat_slerp, quat_product, quat_conjugation):
        setattr(module, fn.__name__, fn)
