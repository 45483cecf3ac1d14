"""VAE decode of sampled latents: the reference's causal VideoVAE and ImageVAE decoders on the
B200 kernels."""
from .image_vae import ImageVAE  # noqa: F401
from .video_vae import VideoVAE  # noqa: F401
