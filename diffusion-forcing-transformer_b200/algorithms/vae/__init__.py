"""VAE decode of sampled latents: the reference's causal VideoVAE decoder on the B200 kernels."""
from .video_vae import VideoVAE  # noqa: F401
