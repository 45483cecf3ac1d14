from .u_vit3d_pose import PoseCondition, UViT3DPose

__all__ = ["UViT3DPose", "PoseCondition"]
