"""TEST INFRASTRUCTURE — CPU restatement (torch fp32) of the decode side of the reference's ImageVAE, the VAE of latent
configurations without temporal compression (`_load_vae`'s default branch, base_pytorch_video_algo.py:541-549; SURVEY.md
§8f rank 1).  Never imported by the product.

Follows, function by function:
  ImageVAE.decode                      algorithms/vae/image_vae/trainer.py:337-340   (post_quant_conv, decoder)
  Decoder.forward                      algorithms/vae/image_vae/model.py:215-245
  ResnetBlock2D.forward                algorithms/vae/common/modules/resnet.py:42-58
  AttnBlock.forward                    algorithms/vae/common/modules/attention.py:58-83 (single head, c^-0.5)
  Upsample.forward                     algorithms/vae/common/modules/updownsample.py:19-24 (nearest x2, conv 3x3)
  Normalize = GroupNorm(32, eps 1e-6)  algorithms/vae/common/modules/normalize.py:4-7;  nonlinearity = x*sigmoid(x)
Pinned against tests/golden/vae_image_decode.npz, produced by executing the reference (oracle/make_goldens_image_vae.py).
Configuration = configurations/algorithm/image_vae.yaml's ddconfig family: no attention at the up levels
(attn_resolutions = []), resamp_with_conv, no tanh.
"""
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F


def image_decoder_param_shapes(ch: int, z_channels: int, embed_dim: int, ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2,
                               out_ch: int = 3) -> List[Tuple[str, Tuple[int, ...]]]:
    """(state-dict key, shape) of every tensor `decode` touches, in the reference's registration order."""
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(name, cin, cout, k):
        out.append((f"{name}.weight", (cout, cin, k, k)))
        out.append((f"{name}.bias", (cout,)))

    def norm(name, c):
        out.append((f"{name}.weight", (c,)))
        out.append((f"{name}.bias", (c,)))

    def resblock(name, cin, cout):
        norm(f"{name}.norm1", cin)
        conv(f"{name}.conv1", cin, cout, 3)
        norm(f"{name}.norm2", cout)
        conv(f"{name}.conv2", cout, cout, 3)
        if cin != cout:
            conv(f"{name}.nin_shortcut", cin, cout, 1)

    L = len(ch_mult)
    block_in = ch * ch_mult[-1]
    conv("decoder.conv_in", z_channels, block_in, 3)
    resblock("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", block_in, block_in, 1)
    resblock("decoder.mid.block_2", block_in, block_in)
    per_level = {}
    for lvl in reversed(range(L)):
        names, block_out = [], ch * ch_mult[lvl]
        for i in range(num_res_blocks + 1):
            names.append((f"decoder.up.{lvl}.block.{i}", block_in, block_out))
            block_in = block_out
        per_level[lvl] = (names, block_in)
    for lvl in range(L):                                   # ModuleList order after `self.up.insert(0, up)`
        names, c = per_level[lvl]
        for n, cin, cout in names:
            resblock(n, cin, cout)
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", c, c, 3)
    norm("decoder.norm_out", ch * ch_mult[0])
    conv("decoder.conv_out", ch * ch_mult[0], out_ch, 3)
    conv("post_quant_conv", embed_dim, z_channels, 1)
    return out


def seeded_image_weights(shapes, seed: int) -> Dict[str, torch.Tensor]:
    """Deterministic stand-in weights: conv weights N(0, 1/fan_in), biases N(0, 0.02^2), norm scales 1 + N(0, 0.1^2),
    norm shifts N(0, 0.1^2); drawn in list order from one CPU generator."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes:
        r = torch.randn(shape, generator=g)
        if len(shape) == 4:
            sd[name] = r / (shape[1] * shape[2] * shape[3]) ** 0.5
        elif "norm" in name.split(".")[-2]:
            sd[name] = 1.0 + 0.1 * r if name.endswith(".weight") else 0.1 * r
        else:
            sd[name] = r * 0.02
    return sd


class ImageVAEDecoderOracle:
    def __init__(self, state_dict: Dict[str, torch.Tensor], ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2):
        self.sd = {k: v.detach().float() for k, v in state_dict.items()}
        self.L, self.nrb = len(ch_mult), num_res_blocks

    def conv(self, name, x, pad):
        return F.conv2d(x, self.sd[f"{name}.weight"], self.sd[f"{name}.bias"], padding=pad)

    def norm(self, name, x):
        return F.group_norm(x, 32, self.sd[f"{name}.weight"], self.sd[f"{name}.bias"], eps=1e-6)

    @staticmethod
    def silu(x):
        return x * torch.sigmoid(x)

    def resblock(self, name, x):                                   # resnet.py:42-58
        h = self.conv(f"{name}.conv1", self.silu(self.norm(f"{name}.norm1", x)), 1)
        h = self.conv(f"{name}.conv2", self.silu(self.norm(f"{name}.norm2", h)), 1)
        if f"{name}.nin_shortcut.weight" in self.sd:
            x = self.conv(f"{name}.nin_shortcut", x, 0)
        return x + h

    def attn(self, name, x):                                       # attention.py:58-83
        h = self.norm(f"{name}.norm", x)
        q, k, v = (self.conv(f"{name}.{n}", h, 0) for n in ("q", "k", "v"))
        b, c, hh, ww = q.shape
        q, k, v = (a.reshape(b, c, hh * ww) for a in (q, k, v))
        w = torch.softmax(torch.bmm(q.permute(0, 2, 1), k) * (int(c) ** -0.5), dim=2)     # [b, query, key]
        o = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
        return x + self.conv(f"{name}.proj_out", o, 0)

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z [n, C, H, W] fp32 -> images [n, 3, 2^(L-1) H, 2^(L-1) W]."""
        h = self.conv("post_quant_conv", z.float(), 0)
        h = self.conv("decoder.conv_in", h, 1)
        h = self.resblock("decoder.mid.block_1", h)
        h = self.attn("decoder.mid.attn_1", h)
        h = self.resblock("decoder.mid.block_2", h)
        for lvl in reversed(range(self.L)):
            for i in range(self.nrb + 1):
                h = self.resblock(f"decoder.up.{lvl}.block.{i}", h)
            if lvl != 0:
                h = self.conv(f"decoder.up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"), 1)
        return self.conv("decoder.conv_out", self.silu(self.norm("decoder.norm_out", h)), 1)
