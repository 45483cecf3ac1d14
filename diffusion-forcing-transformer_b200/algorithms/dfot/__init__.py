from .dfot_video import DFoTVideo

__all__ = ["DFoTVideo"]
