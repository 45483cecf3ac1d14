"""No-GPU: the CPU restatement of the reference's ImageVAE decode side (oracle/image_vae.py) against the fixture
produced by executing the reference (oracle/make_goldens_image_vae.py): same seeded stand-in weights, same latents."""
import json
import os

import numpy as np
import torch

from helpers import GOLDEN
from oracle.image_vae import ImageVAEDecoderOracle, image_decoder_param_shapes, seeded_image_weights


def _case():
    with open(os.path.join(GOLDEN, "vae_image_decode.json")) as f:
        meta = json.load(f)
    return meta, dict(np.load(os.path.join(GOLDEN, "vae_image_decode.npz")))


def test_image_decoder_keys_and_decode_match_reference_fixture():
    meta, arr = _case()
    c = meta["case"]
    dd = c["ddconfig"]
    shapes = image_decoder_param_shapes(dd["ch"], dd["z_channels"], c["embed_dim"], tuple(dd["ch_mult"]), dd["num_res_blocks"])
    assert [n for n, _ in shapes] == meta["keys"]
    dec = ImageVAEDecoderOracle(seeded_image_weights(shapes, c["weight_seed"]), tuple(dd["ch_mult"]), dd["num_res_blocks"])
    images = dec.decode(torch.from_numpy(arr["z"]))
    assert list(images.shape) == meta["image_shape"]
    assert np.abs(images.numpy() - arr["images"]).max() <= 2e-5
