from .dfot_video import DFoTVideo
from .dfot_video_pose import DFoTVideoPose

__all__ = ["DFoTVideo", "DFoTVideoPose"]
