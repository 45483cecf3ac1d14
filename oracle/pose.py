"""ORACLE (test infrastructure): camera-pose conditioning of DFoTVideoPose.

Restates utils/geometry_utils.py (CameraPose.from_vectors :102-117, _normalize_by / normalize_by_first :119-136,
scale_within_bounds :157-168, rays :244-295; Ray.to_pos_encoding :50-81) and
algorithms/dfot/dfot_video_pose.py:64-110 (_process_conditions, fp32).  normalize_by="mean" (:137-155) and the
`temporal` guidance fill-in `replace_with_interpolation` (:170-206) go through roma's quaternion routines, restated in
oracle/roma_restatement.py (roma is absent here; pinned against scipy — see that module)."""
import math

import torch

from . import roma_restatement as roma


def replace_with_interpolation(R: torch.Tensor, Tv: torch.Tensor, mask: torch.Tensor):
    """geometry_utils.py:170-206: poses of masked frames (mask True) <- slerp / lerp between the nearest valid frames,
    constant extension before the first and after the last valid frame; rows that are all valid or all masked are left
    alone.  EVERY rotation (masked or not) comes back through quaternion -> matrix, as in the reference (:205)."""
    q = roma.rotmat_to_unitquat(R)
    Tv = Tv.clone()
    for b in range(mask.shape[0]):
        m = mask[b]
        if not m.any() or m.all():
            continue
        valid = torch.where(~m)[0]
        if valid[0] != 0:
            q[b, : valid[0]] = q[b, valid[0]]
            Tv[b, : valid[0]] = Tv[b, valid[0]]
        if valid[-1] != mask.shape[1] - 1:
            q[b, valid[-1] + 1:] = q[b, valid[-1]]
            Tv[b, valid[-1] + 1:] = Tv[b, valid[-1]]
        for lt, rt in zip(valid[:-1], valid[1:]):
            if rt - lt == 1:
                continue
            steps = torch.linspace(0, 1, int(rt - lt) + 1)
            q[b, lt: rt + 1] = roma.unitquat_slerp(q[b, lt], q[b, rt], steps)
            Tv[b, lt: rt + 1] = torch.lerp(Tv[b, lt], Tv[b, rt], steps.unsqueeze(-1))
    return roma.unitquat_to_rotmat(q), Tv


def ray_encoding(conditions: torch.Tensor, resolution: int, normalize_by: str = "first", bound=None,
                 cond_type: str = "ray_encoding", freq: int = 15, interp_mask=None) -> torch.Tensor:
    """conditions: (B, T, 16) = intrinsics (fx, fy, px, py) + row-major [R | t] (3x4) → (B, T, C, H, W) fp32.
    interp_mask (B, T) bool: frames whose pose is replaced by interpolation first (`temporal` history guidance)."""
    c = conditions.float()
    K, RT = c[..., :4], c[..., 4:].reshape(*c.shape[:2], 3, 4)
    R, Tv = RT[..., :3], RT[..., 3]
    if interp_mask is not None:
        R, Tv = replace_with_interpolation(R, Tv, interp_mask)
    if normalize_by == "first":
        R_ref, T_ref = R[:, 0], Tv[:, 0]
    elif normalize_by == "mean":                                   # geometry_utils.py:137-155
        R_ref = roma.unitquat_to_rotmat(roma.rotmat_to_unitquat(R).mean(dim=1))   # (mean quaternion is NOT re-normalised)
        t_world = torch.einsum("btji,btj->bti", R, Tv).mean(dim=1)
        T_ref = torch.einsum("bij,bj->bi", R_ref, t_world)
    else:
        raise ValueError(f"Unknown camera pose normalization method: {normalize_by}")
    Rrinv = R_ref.transpose(-1, -2)                                # geometry_utils.py:119-126
    R = torch.einsum("btij,bjk->btik", R, Rrinv)
    Tv = Tv - torch.einsum("btij,bj->bti", R, T_ref)              # uses the already re-based rotations
    if bound is not None:                                          # :157-168
        Tv = Tv * (bound / Tv.abs().amax(dim=1, keepdim=True).clamp(min=1e-6))
    # rays (:244-295): pixel centres, meshgrid "xy" → w varies along the last axis
    lin = torch.linspace(0, resolution - 1, resolution, dtype=c.dtype)
    cw = lin[None, :].expand(resolution, resolution) + 0.5
    ch = lin[:, None].expand(resolution, resolution) + 0.5
    Kr = K * resolution
    fx, fy, px, py = (Kr[..., i][..., None, None] for i in range(4))
    x, y = (cw - px) / fx, (ch - py) / fy
    d_cam = torch.stack([x, y, torch.ones_like(x)], dim=-1)        # b t h w 3
    Rinv = R.transpose(-1, -2)
    direction = torch.einsum("btij,bthwj->bthwi", Rinv, d_cam)
    origin = -torch.einsum("btij,btj->bti", Rinv, Tv)
    origin = origin[:, :, None, None, :].expand_as(direction)
    if cond_type == "ray":
        out = torch.cat([origin, direction], dim=-1)
    elif cond_type == "ray_encoding":                              # :50-81
        scale = 2 ** torch.linspace(0, freq - 1, freq, dtype=c.dtype) * math.pi

        def enc(v):
            e = (v[..., None] * scale).flatten(-2)                 # (i s) with the component i slowest
            return torch.sin(torch.cat([e, e + 0.5 * math.pi], dim=-1))
        out = torch.cat([enc(origin), enc(direction)], dim=-1)
    else:
        raise NotImplementedError(cond_type)
    return out.permute(0, 1, 4, 2, 3).contiguous()
