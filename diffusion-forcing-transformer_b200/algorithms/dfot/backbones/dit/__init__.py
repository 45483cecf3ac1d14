from .dit3d import DiT3D

__all__ = ["DiT3D"]
