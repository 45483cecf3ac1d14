"""DFoTVideoPose — camera-pose-conditioned DFoT (RealEstate10K), drop-in for the reference's
algorithms/dfot/dfot_video_pose.py:10-110 on top of the B200 sampler of ``dfot_video.py``.

The reference turns the raw poses into a dense (R, T, 180, H, W) ray-encoding tensor at *every* sampling step
(377 MB per row at 256x256).  Here the per-frame camera table (16 floats) is all that leaves this class: the
backbone turns it into PatchEmbed rows on the fly and caches the resulting per-pixel FiLM modulation once per
window (``UViT3DPose.prepare_pose``).  ``_process_conditions`` keeps the reference signature and, for callers that
really want it, ``ray_encoding()`` materialises the dense tensor with the same kernel.
"""
import math
from typing import Optional

import torch
from torch import Tensor

from dfot_b200 import ops
from . import pose_math
from dfot_b200.config import to_config
from .backbones.u_vit.u_vit3d_pose import PoseCondition
from .dfot_video import DFoTVideo


def ray_freq_scale(n_freq: int = 15) -> Tensor:
    """2^s * pi in fp32, exactly as utils/geometry_utils.py:62-66 builds it (fp32 tensor times a Python scalar)."""
    return 2 ** torch.linspace(0, n_freq - 1, n_freq, dtype=torch.float32) * math.pi


def camera_table(conditions: Tensor, resolution: int, normalize_by: str = "first", bound=None,
                 interp_mask: Optional[Tensor] = None) -> Tensor:
    """(B, T, 16) raw poses (fx, fy, px, py, row-major [R | t]) → (B, T, 16) per-frame camera table
    (fx, fy, px, py in pixels | R^-1 row-major | ray origin -R^-1 t) after the reference's normalisation
    (geometry_utils.py:102-168: left-normalise by frame 0 or by the mean camera, optional bound scaling).  `interp_mask`
    (B, T) bool marks frames whose pose is first rebuilt by interpolation (`temporal` history guidance,
    dfot_video_pose.py:73-81).  16 floats per frame: host-side plumbing, fp32 like the reference (its autocast is
    disabled for this part, dfot_video_pose.py:64-67)."""
    c = conditions.float()
    K, RT = c[..., :4], c[..., 4:].reshape(*c.shape[:2], 3, 4)
    R, t = RT[..., :3], RT[..., 3]
    if interp_mask is not None:
        R, t = pose_math.interpolate_masked_poses(R, t, interp_mask)
    if normalize_by == "first":
        R_ref, t_ref = R[:, 0], t[:, 0]
    elif normalize_by == "mean":   # geometry_utils.py:137-155: mean quaternion (not re-normalised), mean world position
        R_ref = pose_math.quat_to_rotmat(pose_math.rotmat_to_quat(R).mean(dim=1))
        t_ref = torch.einsum("bij,bj->bi", R_ref, torch.einsum("btji,btj->bti", R, t).mean(dim=1))
    else:
        raise ValueError(f"Unknown camera pose normalization method: {normalize_by}")
    R = torch.einsum("btij,bjk->btik", R, R_ref.transpose(-1, -2))     # (the reference "inverts" by transposing, :124)
    t = t - torch.einsum("btij,bj->bti", R, t_ref)
    if bound is not None:
        t = t * (bound / t.abs().amax(dim=1, keepdim=True).clamp(min=1e-6))
    R_inv = R.transpose(-1, -2)   # geometry_utils.py:283: the transpose, also for the scaled rotations of "mean"
    origin = -torch.einsum("btij,btj->bti", R_inv, t)
    return torch.cat([K * resolution, R_inv.reshape(*c.shape[:2], 9), origin], dim=-1).contiguous()


class DFoTVideoPose(DFoTVideo):
    def __init__(self, cfg):
        cfg = to_config(cfg)
        self.camera_pose_conditioning = cfg.camera_pose_conditioning
        self.conditioning_type = cfg.camera_pose_conditioning.type
        self._check_cfg(cfg)
        self._update_backbone_cfg(cfg)
        super().__init__(cfg)

    def _check_cfg(self, cfg):
        if cfg.backbone.name not in {"dit3d_pose", "u_vit3d_pose"}:
            raise ValueError("DiffusionForcingVideo3D only supports backbone 'dit3d_pose' or 'u_vit3d_pose', "
                             f"got {cfg.backbone.name}")
        if cfg.backbone.name == "u_vit3d_pose" and self.conditioning_type == "global":
            raise ValueError("Global camera pose conditioning is not supported for U-ViT3DPose")
        if cfg.backbone.name == "dit3d_pose":
            raise NotImplementedError("backbone `dit3d_pose` is not implemented by dfot_b200 (the RE10K configs use "
                                      "u_vit3d_pose)")
        if self.conditioning_type != "ray_encoding":
            raise NotImplementedError(f"camera_pose_conditioning.type={self.conditioning_type} is not implemented by "
                                      "dfot_b200 (RE10K uses ray_encoding)")

    def _update_backbone_cfg(self, cfg):
        dims = {"global": 12, "ray": 6, "plucker": 6, "ray_encoding": 180}
        if self.conditioning_type not in dims:
            raise ValueError(f"Unknown camera pose conditioning type: {self.conditioning_type}")
        cfg.backbone.conditioning.dim = dims[self.conditioning_type]

    # ------------------------------------------------------------------ conditioning
    @torch.no_grad()
    def _process_conditions(self, conditions: Tensor, noise_levels: Optional[Tensor] = None) -> PoseCondition:
        """Reference signature; returns the lightweight handle the UViT3DPose backbone consumes."""
        return self._window_conditions(conditions, 1)

    def _conditions_follow_levels(self) -> bool:
        """`temporal` history guidance: frames at the top noise level are 'fully masked' and their camera poses are
        rebuilt by interpolation (dfot_video_pose.py:73-81) — the conditioning then depends on the step's noise levels."""
        return self.cfg.tasks.prediction.history_guidance.name == "temporal"

    def _window_conditions(self, conditions: Optional[Tensor], nfe: int, levels_from=None):
        if conditions is None:
            return None
        cp = self.camera_pose_conditioning
        conditions = conditions.to(self.device)
        if self._conditions_follow_levels() and levels_from is not None:
            # one camera table per backbone row: rows are ordered (b nfe) like the reference's repeat (dfot_video.py:733-738)
            mask = torch.as_tensor(levels_from == self.timesteps - 1, device=conditions.device)
            rows = conditions.repeat_interleave(nfe, dim=0)
            cams = camera_table(rows, self.x_shape[1], cp.normalize_by, cp.bound, interp_mask=mask)
            return PoseCondition(cams, list(range(rows.shape[0])))
        cams = camera_table(conditions, self.x_shape[1], cp.normalize_by, cp.bound)
        rows = [b for b in range(conditions.shape[0]) for _ in range(nfe)]
        return PoseCondition(cams, rows)

    @torch.no_grad()
    def ray_encoding(self, conditions: Tensor) -> Tensor:
        """The reference's dense (B, T, 180, H, W) fp32-valued ray encoding (rounded to bf16 by the kernel)."""
        cp = self.camera_pose_conditioning
        res, B, T = self.x_shape[1], conditions.shape[0], conditions.shape[1]
        cams = camera_table(conditions.to(self.device), res, cp.normalize_by, cp.bound)
        out = torch.empty((B * T * res * res, 180), dtype=torch.bfloat16, device=self.device)
        ops.pose_ray_patches(cams.reshape(B * T, 16).contiguous(), ray_freq_scale().to(self.device), out, B * T, res, 1)
        return out.view(B, T, res, res, 180).permute(0, 1, 4, 2, 3).float()
