"""Quaternion arithmetic of the camera-pose path (host side, fp32, a few hundred floats per window).

The reference does this with roma 1.5.2.1 (utils/geometry_utils.py:141-143 `normalize_by_mean`, :172-205
`replace_with_interpolation`): rotation matrix <-> unit quaternion (XYZW) and slerp.  roma's mappings follow SciPy's
`Rotation` (largest-of-diagonal-or-trace construction, `as_matrix` formula, rotation-vector exponential with Taylor
branches below 1e-3 rad); the same arithmetic is written here on batched tensors, without roma.  One behaviour matters for
parity: `quat_to_rotmat` does not normalise, so the arithmetic mean of unit quaternions gives a rotation scaled by |q|^2,
exactly what the reference's `normalize_by_mean` feeds into `_normalize_by` (geometry_utils.py:142-155).
"""
import torch
from torch import Tensor


def rotmat_to_quat(R: Tensor) -> Tensor:
    """(..., 3, 3) -> (..., 4) unit quaternion, XYZW.  Branch on argmax(R00, R11, R22, trace) like SciPy / roma; all four
    candidate constructions are evaluated and the chosen one is gathered (no data-dependent indexing)."""
    m = R.reshape(-1, 3, 3)
    d0, d1, d2 = m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]
    tr = d0 + d1 + d2
    choice = torch.stack([d0, d1, d2, tr], dim=1).argmax(dim=1)
    cand = []
    for i in range(3):
        j, k = (i + 1) % 3, (i + 2) % 3
        q = torch.empty((m.shape[0], 4), dtype=m.dtype, device=m.device)
        q[:, i] = 1 - tr + 2 * m[:, i, i]
        q[:, j] = m[:, j, i] + m[:, i, j]
        q[:, k] = m[:, k, i] + m[:, i, k]
        q[:, 3] = m[:, k, j] - m[:, j, k]
        cand.append(q)
    cand.append(torch.stack([m[:, 2, 1] - m[:, 1, 2], m[:, 0, 2] - m[:, 2, 0], m[:, 1, 0] - m[:, 0, 1], 1 + tr], dim=1))
    q = torch.stack(cand, dim=1)[torch.arange(m.shape[0], device=m.device), choice]
    q = q / torch.norm(q, dim=1, keepdim=True)
    return q.reshape(*R.shape[:-2], 4)


def quat_to_rotmat(q: Tensor) -> Tensor:
    """(..., 4) XYZW -> (..., 3, 3), unit-quaternion formula WITHOUT normalisation (see the module docstring)."""
    x, y, z, w = q.unbind(-1)
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    rows = [x2 - y2 - z2 + w2, 2 * (xy - zw), 2 * (xz + yw),
            2 * (xy + zw), -x2 + y2 - z2 + w2, 2 * (yz - xw),
            2 * (xz - yw), 2 * (yz + xw), -x2 - y2 + z2 + w2]
    return torch.stack(rows, dim=-1).reshape(*q.shape[:-1], 3, 3)


def quat_mul(p: Tensor, q: Tensor) -> Tensor:
    """Hamilton product, XYZW."""
    pv, pw, qv, qw = p[..., :3], p[..., 3:], q[..., :3], q[..., 3:]
    return torch.cat([pw * qv + qw * pv + torch.cross(pv, qv, dim=-1), pw * qw - (pv * qv).sum(-1, keepdim=True)], dim=-1)


def quat_slerp(q0: Tensor, q1: Tensor, steps: Tensor) -> Tensor:
    """q0 * exp(s * log(q0^-1 q1)) along the shorter arc for every s in `steps` (n,) -> (n, 4); q0, q1: (4,)."""
    rel = quat_mul(torch.cat([-q0[:3], q0[3:]]), q1)
    if rel[3] < 0:                                   # shortest arc: w >= 0
        rel = -rel
    vn = torch.norm(rel[:3])
    half = torch.atan2(vn, rel[3])
    angle = 2 * half
    if torch.abs(angle) <= 1e-3:
        scale = 2 + angle ** 2 / 12 + 7 * angle ** 4 / 2880
    else:
        scale = angle / torch.sin(half)
    rotvec = steps[:, None] * (scale * rel[:3])[None, :]             # (n, 3)
    norms = torch.norm(rotvec, dim=1)
    small = norms <= 1e-3
    safe = torch.where(small, torch.ones_like(norms), norms)
    s = torch.where(small, 0.5 - norms ** 2 / 48 + norms ** 4 / 3840, torch.sin(safe / 2) / safe)
    rot = torch.cat([s[:, None] * rotvec, torch.cos(norms / 2)[:, None]], dim=1)
    return quat_mul(q0[None, :].expand(steps.shape[0], 4), rot)


def interpolate_masked_poses(R: Tensor, t: Tensor, mask: Tensor):
    """geometry_utils.py:170-206.  R (B, T, 3, 3), t (B, T, 3), mask (B, T) bool — True = the frame is fully masked out
    by `temporal` history guidance and its pose is rebuilt from the nearest unmasked frames (slerp / lerp between them,
    constant extension at the ends).  Rows without masked frames, or with nothing left, keep their poses; like the
    reference, every rotation of the batch takes the matrix -> quaternion -> matrix round trip."""
    q = rotmat_to_quat(R)
    t = t.clone()
    mask_h = mask.detach().cpu()
    T = mask_h.shape[1]
    for b in range(mask_h.shape[0]):
        m = mask_h[b]
        if not bool(m.any()) or bool(m.all()):
            continue
        valid = torch.where(~m)[0].tolist()
        if valid[0] != 0:
            q[b, : valid[0]] = q[b, valid[0]]
            t[b, : valid[0]] = t[b, valid[0]]
        if valid[-1] != T - 1:
            q[b, valid[-1] + 1:] = q[b, valid[-1]]
            t[b, valid[-1] + 1:] = t[b, valid[-1]]
        for lt, rt in zip(valid[:-1], valid[1:]):
            if rt - lt == 1:
                continue
            steps = torch.linspace(0, 1, rt - lt + 1, device=q.device, dtype=q.dtype)
            q[b, lt: rt + 1] = quat_slerp(q[b, lt], q[b, rt], steps)
            t[b, lt: rt + 1] = torch.lerp(t[b, lt], t[b, rt], steps[:, None])
    return quat_to_rotmat(q), t
