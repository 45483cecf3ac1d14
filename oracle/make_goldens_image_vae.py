"""TEST INFRASTRUCTURE — tests/golden/vae_image_decode.npz by EXECUTING the reference's ImageVAE decode side (authoring
container only, needs /root/reference):   python -m oracle.make_goldens_image_vae
The reference's `Decoder` (algorithms/vae/image_vae/model.py) and a `post_quant_conv` are built exactly as
`ImageVAE.__init__` builds them (trainer.py:287-296) — the trainer module itself drags in Lightning and the LPIPS loss —
and run on seeded stand-in weights (oracle.image_vae.seeded_image_weights, regenerated identically by the tests)."""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.image_vae import image_decoder_param_shapes, seeded_image_weights  # noqa: E402

CASE = dict(ddconfig=dict(double_z=True, z_channels=4, resolution=32, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2],
                          num_res_blocks=2, attn_resolutions=[], dropout=0.0),
            embed_dim=4, weight_seed=21, data_seed=22, batch=3)


def import_reference_decoder():
    ref_shim.install()
    for pkg in ["algorithms.vae", "algorithms.vae.common", "algorithms.vae.image_vae"]:   # path-only: skip heavy __init__
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(ref_shim.REF, *pkg.split("."))]
        sys.modules[pkg] = m
    from algorithms.vae.image_vae.model import Decoder
    return Decoder


class _RefDecodeSide(torch.nn.Module):
    """the decode-side attributes of ImageVAE, same names -> same state-dict keys"""

    def __init__(self, Decoder, ddconfig, embed_dim):
        super().__init__()
        self.decoder = Decoder(**ddconfig)
        self.post_quant_conv = torch.nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)

    def decode(self, z):                                    # trainer.py:337-340
        return self.decoder(self.post_quant_conv(z))


def main():
    Decoder = import_reference_decoder()
    c = CASE
    dd = c["ddconfig"]
    torch.manual_seed(0)
    vae = _RefDecodeSide(Decoder, dd, c["embed_dim"]).eval()
    shapes = image_decoder_param_shapes(dd["ch"], dd["z_channels"], c["embed_dim"], tuple(dd["ch_mult"]), dd["num_res_blocks"])
    ref_keys = list(vae.state_dict().keys())
    assert [n for n, _ in shapes] == ref_keys, "decoder state-dict keys / order differ from the reference"
    assert all(tuple(vae.state_dict()[n].shape) == s for n, s in shapes)
    sd = seeded_image_weights(shapes, c["weight_seed"])
    vae.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(c["data_seed"])
    lat = dd["resolution"] // 2 ** (len(dd["ch_mult"]) - 1)
    z = torch.randn((c["batch"], c["embed_dim"], lat, lat), generator=g)
    with torch.no_grad():
        images = vae.decode(z)
    out = os.path.join(ROOT, "tests", "golden")
    np.savez_compressed(os.path.join(out, "vae_image_decode.npz"), z=z.numpy(), images=images.numpy())
    with open(os.path.join(out, "vae_image_decode.json"), "w") as f:
        json.dump(dict(case=c, keys=ref_keys, image_shape=list(images.shape)), f, indent=1)
    print("images", tuple(images.shape), "abs max", float(images.abs().max()), "keys", len(ref_keys))


if __name__ == "__main__":
    main()
