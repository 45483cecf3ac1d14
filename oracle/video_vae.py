"""TEST INFRASTRUCTURE — CPU restatement (torch fp32) of the decoder of the reference's causal VideoVAE, the step that
follows the sampling path for latent configurations.  Never imported by the product.

Follows, function by function:
  VideoVAE.decode / _decode            algorithms/vae/video_vae/model.py:449-481   (post_quant_conv, decoder, tail slice)
  Decoder.forward                      algorithms/vae/video_vae/model.py:252-270
  PaddedConv3D.forward (causal)        algorithms/vae/common/modules/conv.py:98-108 (first frame repeated kt-1 times)
  ResnetBlock3D.forward                algorithms/vae/common/modules/resnet.py:93-109
  AttnBlock3D.forward                  algorithms/vae/common/modules/attention.py:115-156 (per-frame, single head, c^-0.5)
  SpatialUpsample2x.forward            algorithms/vae/common/modules/updownsample.py:73-80  (nearest x2, conv (1,3,3))
  Spatial2xTime2x3DUpsample.forward    algorithms/vae/common/modules/updownsample.py:131-147 (first frame spatial-only,
                                       the rest trilinear x(2,2,2), align_corners=False; then conv 3x3x3)
  Normalize = GroupNorm(32, eps 1e-6)  algorithms/vae/common/modules/normalize.py:4-7;  nonlinearity = x*sigmoid(x)
Pinned against tests/golden/vae_video_decode.npz, produced by executing the reference (oracle/make_goldens_vae.py).
Configuration = the VideoVAE constructor defaults (model.py:282-342): decoder levels (reversed) 3: ResnetBlock3D x3 +
Spatial2xTime2x3DUpsample, 2: same, 1: x3 + SpatialUpsample2x, 0: x3; mid = ResnetBlock3D, AttnBlock3D, ResnetBlock3D;
conv_in / conv_out / post_quant_conv are PaddedConv3D; no attention at the up levels (attn_resolutions = ()).
"""
import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F


def decoder_param_shapes(hidden_size: int, z_channels: int, embed_dim: int, mult=(1, 2, 4, 4),
                         num_res_blocks: int = 2) -> List[Tuple[str, Tuple[int, ...]]]:
    """(state-dict key, shape) of every tensor `decode` touches, in the reference's registration order."""
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(name, cin, cout, k):
        out.append((f"{name}.conv.weight", (cout, cin, *k)))
        out.append((f"{name}.conv.bias", (cout,)))

    def norm(name, c):
        out.append((f"{name}.weight", (c,)))
        out.append((f"{name}.bias", (c,)))

    def resblock(name, cin, cout):
        norm(f"{name}.norm1", cin)
        conv(f"{name}.conv1", cin, cout, (3, 3, 3))
        norm(f"{name}.norm2", cout)
        conv(f"{name}.conv2", cout, cout, (3, 3, 3))
        if cin != cout:
            conv(f"{name}.nin_shortcut", cin, cout, (1, 1, 1))

    L = len(mult)
    block_in = hidden_size * mult[-1]
    conv("decoder.conv_in", z_channels, block_in, (3, 3, 3))
    resblock("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", block_in, block_in, (1, 1, 1))
    resblock("decoder.mid.block_2", block_in, block_in)
    per_level = {}
    for lvl in reversed(range(L)):
        names = []
        block_out = hidden_size * mult[lvl]
        for i in range(num_res_blocks + 1):
            names.append((f"decoder.up.{lvl}.block.{i}", block_in, block_out))
            block_in = block_out
        per_level[lvl] = (names, block_in)
    # nn.ModuleList order after `self.up.insert(0, up)`: level 0 first
    for lvl in range(L):
        names, ch = per_level[lvl]
        for n, cin, cout in names:
            resblock(n, cin, cout)
        if lvl >= 1:
            conv(f"decoder.up.{lvl}.upsample.conv", ch, ch, (1, 3, 3) if lvl == 1 else (3, 3, 3))
    norm("decoder.norm_out", hidden_size * mult[0])
    conv("decoder.conv_out", hidden_size * mult[0], 3, (3, 3, 3))
    conv("post_quant_conv", embed_dim, z_channels, (1, 1, 1))
    return out


def seeded_weights(shapes, seed: int) -> Dict[str, torch.Tensor]:
    """Deterministic stand-in weights (the reference ships none): conv weights N(0, 1/fan_in), biases N(0, 0.02^2),
    norm scales 1 + N(0, 0.1^2), norm shifts N(0, 0.1^2); drawn in list order from one CPU generator."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes:
        r = torch.randn(shape, generator=g)
        if name.endswith("conv.weight"):
            sd[name] = r / math.sqrt(shape[1] * shape[2] * shape[3] * shape[4])
        elif name.endswith("conv.bias"):
            sd[name] = r * 0.02
        elif name.endswith(".weight"):
            sd[name] = 1.0 + 0.1 * r
        else:
            sd[name] = 0.1 * r
    return sd


class VideoVAEDecoderOracle:
    def __init__(self, state_dict: Dict[str, torch.Tensor], mult=(1, 2, 4, 4), num_res_blocks: int = 2):
        self.sd = {k: v.detach().float() for k, v in state_dict.items()}
        self.L, self.nrb = len(mult), num_res_blocks

    # conv.py:98-108 — causal: the first frame is repeated (kt - 1) times in front, no temporal padding otherwise
    def conv(self, name, x, spatial_pad):
        w, b = self.sd[f"{name}.conv.weight"], self.sd[f"{name}.conv.bias"]
        kt = w.shape[2]
        if kt > 1:
            x = torch.cat([x[:, :, :1].repeat(1, 1, kt - 1, 1, 1), x], 2)
        return F.conv3d(x, w, b, padding=(0, spatial_pad, spatial_pad))

    def norm(self, name, x):
        return F.group_norm(x, 32, self.sd[f"{name}.weight"], self.sd[f"{name}.bias"], eps=1e-6)

    @staticmethod
    def silu(x):
        return x * torch.sigmoid(x)

    def resblock(self, name, x):                                   # resnet.py:93-109
        h = self.conv(f"{name}.conv1", self.silu(self.norm(f"{name}.norm1", x)), 1)
        h = self.conv(f"{name}.conv2", self.silu(self.norm(f"{name}.norm2", h)), 1)
        if f"{name}.nin_shortcut.conv.weight" in self.sd:
            x = self.conv(f"{name}.nin_shortcut", x, 0)
        return x + h

    def attn(self, name, x):                                       # attention.py:115-156
        h = self.norm(f"{name}.norm", x)
        q, k, v = (self.conv(f"{name}.{n}", h, 0) for n in ("q", "k", "v"))
        b, c, t, hh, ww = q.shape
        flat = lambda a: a.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh * ww)
        q, k, v = flat(q), flat(k), flat(v)
        w = torch.softmax(torch.bmm(q.transpose(1, 2), k) * (int(c) ** -0.5), dim=2)      # [bt, query, key]
        o = torch.bmm(v, w.transpose(1, 2)).reshape(b, t, c, hh, ww).permute(0, 2, 1, 3, 4)
        return x + self.conv(f"{name}.proj_out", o, 0)

    def upsample(self, lvl, x):
        name = f"decoder.up.{lvl}.upsample.conv"
        if lvl == 1:                                               # updownsample.py:73-80
            b, c, t, h, w = x.shape
            x = F.interpolate(x.reshape(b, c * t, h, w), scale_factor=(2, 2), mode="nearest").reshape(b, c, t, 2 * h, 2 * w)
            return self.conv(name, x, 1)
        if x.shape[2] > 1:                                         # updownsample.py:131-147
            first = F.interpolate(x[:, :, :1], scale_factor=(1, 2, 2), mode="trilinear")
            rest = F.interpolate(x[:, :, 1:], scale_factor=(2, 2, 2), mode="trilinear")
            x = torch.cat([first, rest], 2)
        else:
            x = F.interpolate(x, scale_factor=(1, 2, 2), mode="trilinear")
        return self.conv(name, x, 1)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, desired_length=None) -> torch.Tensor:
        """z [B, C, T, H, W] fp32 -> video [B, 3, 1 + 4 (T - 1), 8 H, 8 W] (model.py:449-481, 252-270)."""
        h = self.conv("post_quant_conv", z.float(), 0)
        h = self.conv("decoder.conv_in", h, 1)
        h = self.resblock("decoder.mid.block_1", h)
        h = self.attn("decoder.mid.attn_1", h)
        h = self.resblock("decoder.mid.block_2", h)
        for lvl in reversed(range(self.L)):
            for i in range(self.nrb + 1):
                h = self.resblock(f"decoder.up.{lvl}.block.{i}", h)
            if lvl >= 1:
                h = self.upsample(lvl, h)
        h = self.conv("decoder.conv_out", self.silu(self.norm("decoder.norm_out", h)), 1)
        return h if desired_length is None else h[:, :, -desired_length:]
