from .dit.dit3d import DiT3D
from .u_vit.u_vit3d_pose import PoseCondition, UViT3DPose

__all__ = ["DiT3D", "UViT3DPose", "PoseCondition"]
