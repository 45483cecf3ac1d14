"""ORACLE (test infrastructure): camera-pose conditioning of DFoTVideoPose.

Restates utils/geometry_utils.py (CameraPose.from_vectors :102-117, _normalize_by / normalize_by_first :119-136,
scale_within_bounds :157-168, rays :244-295; Ray.to_pos_encoding :50-81) and
algorithms/dfot/dfot_video_pose.py:64-110 (_process_conditions, fp32).  normalize_by="mean" and the `temporal`
guidance interpolation need roma (absent here) and are out of scope."""
import math

import torch


def ray_encoding(conditions: torch.Tensor, resolution: int, normalize_by: str = "first", bound=None,
                 cond_type: str = "ray_encoding", freq: int = 15) -> torch.Tensor:
    """conditions: (B, T, 16) = intrinsics (fx, fy, px, py) + row-major [R | t] (3x4) → (B, T, C, H, W) fp32."""
    c = conditions.float()
    K, RT = c[..., :4], c[..., 4:].reshape(*c.shape[:2], 3, 4)
    R, Tv = RT[..., :3], RT[..., 3]
    if normalize_by != "first":
        raise NotImplementedError("only normalize_by='first' is covered by the oracle")
    R0inv = R[:, 0].transpose(-1, -2)                              # geometry_utils.py:119-136
    R = torch.einsum("btij,bjk->btik", R, R0inv)
    Tv = Tv - torch.einsum("btij,bj->bti", R, Tv[:, 0])          # uses the already re-based rotations
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
