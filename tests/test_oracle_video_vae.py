"""No-GPU: the CPU restatement of the reference's causal VideoVAE decoder (oracle/video_vae.py — SURVEY.md §8f rank 1,
the step after the sampling path for latent configurations) against the fixture produced by executing the reference
(oracle/make_goldens_vae.py): same seeded stand-in weights, same latent."""
import json
import os

import numpy as np
import torch

from helpers import GOLDEN
from oracle.video_vae import VideoVAEDecoderOracle, decoder_param_shapes, seeded_weights


def _case():
    with open(os.path.join(GOLDEN, "vae_video_decode.json")) as f:
        meta = json.load(f)
    return meta, dict(np.load(os.path.join(GOLDEN, "vae_video_decode.npz")))


def test_decoder_state_dict_keys_match_reference():
    meta, _ = _case()
    c = meta["case"]
    shapes = decoder_param_shapes(c["hidden_size"], c["z_channels"], c["embed_dim"], tuple(c["hidden_size_mult"]))
    assert [n for n, _ in shapes] == meta["keys"]


def test_decode_matches_reference_fixture():
    meta, arr = _case()
    c = meta["case"]
    shapes = decoder_param_shapes(c["hidden_size"], c["z_channels"], c["embed_dim"], tuple(c["hidden_size_mult"]))
    dec = VideoVAEDecoderOracle(seeded_weights(shapes, c["weight_seed"]), tuple(c["hidden_size_mult"]))
    z = torch.from_numpy(arr["z"])
    video = dec.decode(z, c["temporal_length"])
    assert list(video.shape) == meta["video_shape"]                      # 3 latent frames -> 1 + 4*2 = 9 frames, x8 spatial
    assert np.abs(video.numpy() - arr["video"]).max() <= 2e-5
    short = dec.decode(z[:, :, :2], 5)                                   # a shorter latent clip (2 latent frames -> 5)
    assert np.abs(short.numpy() - arr["short"]).max() <= 2e-5
