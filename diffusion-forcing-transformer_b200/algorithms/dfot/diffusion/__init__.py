from .discrete_diffusion import DiscreteDiffusion
from .continuous_diffusion import ContinuousDiffusion

__all__ = ["DiscreteDiffusion", "ContinuousDiffusion"]
