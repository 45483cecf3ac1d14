"""TEST INFRASTRUCTURE — tests/golden/vae_video_decode.npz by EXECUTING the reference's VideoVAE (authoring
container only, needs /root/reference):   python -m oracle.make_goldens_vae
The reference holds no fixture for the VAE; the decoder is run on seeded stand-in weights (oracle.video_vae.seeded_weights,
regenerated identically by the tests, so only the latent and the decoded video are stored)."""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.video_vae import decoder_param_shapes, seeded_weights  # noqa: E402

CASE = dict(hidden_size=32, z_channels=4, embed_dim=4, hidden_size_mult=[1, 2, 2, 2], resolution=32, temporal_length=9,
            weight_seed=11, data_seed=12, batch=2, latent_frames=3)


def import_reference_vae():
    ref_shim.install()
    for pkg in ["algorithms.vae", "algorithms.vae.common", "algorithms.vae.video_vae"]:   # path-only: skip heavy __init__
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(ref_shim.REF, *pkg.split("."))]
        sys.modules[pkg] = m
    ck = types.ModuleType("utils.ckpt_utils")
    for n in ["is_wandb_run_path", "is_hf_path", "wandb_to_local_path", "download_pretrained"]:
        setattr(ck, n, lambda *a, **k: False)
    sys.modules["utils.ckpt_utils"] = ck
    from algorithms.vae.video_vae.model import VideoVAE
    return VideoVAE


def main():
    VideoVAE = import_reference_vae()
    c = CASE
    torch.manual_seed(0)
    vae = VideoVAE(hidden_size=c["hidden_size"], z_channels=c["z_channels"], embed_dim=c["embed_dim"],
                   hidden_size_mult=tuple(c["hidden_size_mult"]), resolution=c["resolution"],
                   temporal_length=c["temporal_length"]).eval()
    shapes = decoder_param_shapes(c["hidden_size"], c["z_channels"], c["embed_dim"], tuple(c["hidden_size_mult"]))
    ref_keys = [k for k in vae.state_dict() if k.startswith(("decoder.", "post_quant_conv."))]
    assert [n for n, _ in shapes] == ref_keys, "decoder state-dict keys / order differ from the reference"
    sd = seeded_weights(shapes, c["weight_seed"])
    missing, unexpected = vae.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("encoder.", "quant_conv.")) for k in missing)
    g = torch.Generator().manual_seed(c["data_seed"])
    z = torch.randn((c["batch"], c["embed_dim"], c["latent_frames"], c["resolution"] // 8, c["resolution"] // 8), generator=g)
    with torch.no_grad():
        video = vae.decode(z, c["temporal_length"])
        short = vae.decode(z[:, :, :2], 5)
    out = os.path.join(ROOT, "tests", "golden")
    np.savez_compressed(os.path.join(out, "vae_video_decode.npz"), z=z.numpy(), video=video.numpy(), short=short.numpy())
    with open(os.path.join(out, "vae_video_decode.json"), "w") as f:
        json.dump(dict(case=c, keys=ref_keys, video_shape=list(video.shape)), f, indent=1)
    print("video", tuple(video.shape), "abs max", float(video.abs().max()), "keys", len(ref_keys))


if __name__ == "__main__":
    main()
