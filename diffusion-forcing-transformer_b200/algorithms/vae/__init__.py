"""VAE decode of sampled latents: the reference's causal VideoVAE, its ImageVAE and the DC-AE image
autoencoder of the DMLab / Minecraft configurations — decoders on the B200 kernels."""
from .dc_ae import MyAutoencoderDC  # noqa: F401
from .image_vae import ImageVAE  # noqa: F401
from .video_vae import VideoVAE  # noqa: F401
