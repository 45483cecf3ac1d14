"""Decoder of the reference's ImageVAE on the B200 kernels — the VAE of latent configurations without temporal
compression (`_load_vae`'s default branch, algorithms/common/base_pytorch_video_algo.py:541-549).
Mirrors, on the decode side only,

  ImageVAE.__init__ / decode / from_pretrained   algorithms/vae/image_vae/trainer.py:281-340
  Decoder.forward                                algorithms/vae/image_vae/model.py:128-245
  ResnetBlock2D / AttnBlock / Upsample           algorithms/vae/common/modules/{resnet.py:8-58, attention.py:39-83, updownsample.py:10-24}

with the same `cfg` (ddconfig + embed_dim) and the same state-dict keys (`decoder.*`, `post_quant_conv.*`).  It is the
2-D case of the VideoVAE machinery (video_vae.py): images are clips of one frame with no pad slots, every convolution is
the kt = 1 case of the implicit-GEMM kernel, every level but the last is upsampled nearest x2.  CUDA only.
"""
import math
from typing import Dict

import torch
from torch import Tensor, nn

from ... import ops
from ...checkpoint_io import load_checkpoint_file
from ...config import to_config
from .video_vae import _DecoderOnKernels, _register

# configurations/algorithm/image_vae.yaml — used by the reference for checkpoints that carry no `cfg` (trainer.py:313-318)
_DEFAULT_CFG = dict(embed_dim=4, ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                                               ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0))


def _image_decoder_params(ch, z_channels, embed_dim, ch_mult, num_res_blocks, out_ch):
    """(key, shape) of the decode-side parameters in the reference's registration order (model.py:128-213)."""
    out = []

    def conv(name, cin, cout, k):
        out.extend([(f"{name}.weight", (cout, cin, k, k)), (f"{name}.bias", (cout,))])

    def norm(name, c):
        out.extend([(f"{name}.weight", (c,)), (f"{name}.bias", (c,))])

    def res(name, cin, cout):
        norm(f"{name}.norm1", cin)
        conv(f"{name}.conv1", cin, cout, 3)
        norm(f"{name}.norm2", cout)
        conv(f"{name}.conv2", cout, cout, 3)
        if cin != cout:
            conv(f"{name}.nin_shortcut", cin, cout, 1)

    L = len(ch_mult)
    c = ch * ch_mult[-1]
    conv("decoder.conv_in", z_channels, c, 3)
    res("decoder.mid.block_1", c, c)
    norm("decoder.mid.attn_1.norm", c)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", c, c, 1)
    res("decoder.mid.block_2", c, c)
    levels = {}
    for lvl in reversed(range(L)):
        cout, blocks = ch * ch_mult[lvl], []
        for i in range(num_res_blocks + 1):
            blocks.append((f"decoder.up.{lvl}.block.{i}", c, cout))
            c = cout
        levels[lvl] = (blocks, c)
    for lvl in range(L):
        blocks, c = levels[lvl]
        for b in blocks:
            res(*b)
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", c, c, 3)
    norm("decoder.norm_out", ch * ch_mult[0])
    conv("decoder.conv_out", ch * ch_mult[0], out_ch, 3)
    conv("post_quant_conv", embed_dim, z_channels, 1)
    return out


class ImageVAE(_DecoderOnKernels):
    pad = 0

    def _conv_key(self, name: str) -> str:
        return name                                           # plain nn.Conv2d: `<module>.weight`

    def __init__(self, cfg):
        super().__init__()
        cfg = to_config(cfg)
        dd = cfg.ddconfig
        unsupported = [k for k, bad in (("attn_resolutions", bool(list(dd.get("attn_resolutions") or []))),
                                        ("resamp_with_conv", dd.get("resamp_with_conv", True) is not True),
                                        ("tanh_out", bool(dd.get("tanh_out", False))),
                                        ("give_pre_end", bool(dd.get("give_pre_end", False))),
                                        ("use_linear_attn", bool(dd.get("use_linear_attn", False))),
                                        ("attn_type", dd.get("attn_type", "vanilla") != "vanilla")) if bad]
        if unsupported:
            raise NotImplementedError(f"ImageVAE: ddconfig options {unsupported} are not built by dfot_b200")
        if dd.ch % 32 or dd.out_ch > 8:
            raise ValueError("ImageVAE: ch must be a multiple of 32 (GroupNorm groups) and out_ch <= 8")
        self.hidden_size, self.mult, self.nrb = dd.ch, tuple(dd.ch_mult), dd.num_res_blocks
        self.z_channels, self.embed_dim, self.out_ch = dd.z_channels, cfg.embed_dim, dd.out_ch
        g = torch.Generator().manual_seed(0)
        for key, shape in _image_decoder_params(dd.ch, dd.z_channels, cfg.embed_dim, self.mult, self.nrb, dd.out_ch):
            if len(shape) == 4:
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[1] * shape[2] * shape[3])
            else:
                t = torch.ones(shape) if ("norm" in key and key.endswith(".weight")) else torch.zeros(shape)
            _register(self, key, nn.Parameter(t, requires_grad=False))
        self._init_runtime()

    # ------------------------------------------------------------------ checkpoint (trainer.py:298-328)
    @classmethod
    def from_pretrained(cls, path: str, **kwargs) -> "ImageVAE":
        if path.startswith("diffuser:"):
            raise NotImplementedError("ImageVAE: diffusers AutoencoderKL checkpoints are outside the scope of dfot_b200")
        ckpt: Dict = load_checkpoint_file(path)
        model = cls(ckpt.get("cfg", _DEFAULT_CFG))
        sd = {k: v for k, v in ckpt["state_dict"].items() if not k.startswith("loss")}
        own = [n for n, _ in model.named_parameters()]
        missing = [n for n in own if n not in sd]
        if missing:
            raise RuntimeError(f"ImageVAE.from_pretrained: checkpoint lacks decoder tensors {missing[:4]} ...")
        model.load_state_dict(sd)
        return model

    # ------------------------------------------------------------------ decode (trainer.py:337-340, model.py:215-245)
    def _upsample(self, lvl: int, h: Tensor, B, T, H, W, ch):
        if lvl == 0:
            return None
        u16 = self._buf("a16", (B, T, 2 * H, 2 * W, ch), torch.bfloat16, h.device)
        ops.upsample2x_nearest_bf16(h, u16, B * T, H, W, ch)
        return u16, T

    @torch.no_grad()
    def decode(self, z: Tensor) -> Tensor:
        """z [n, embed_dim, H, W] -> images [n, out_ch, 2^(L-1) H, 2^(L-1) W] fp32."""
        ops.require_cuda(z.device, "ImageVAE.decode")
        n, Cz, H, W = z.shape
        if Cz != self.embed_dim:
            raise ValueError(f"ImageVAE.decode: expected {self.embed_dim} latent channels, got {Cz}")
        out, _ = self._run_decoder(z.permute(0, 2, 3, 1).unsqueeze(1), n, 1, H, W, Cz)
        return out[:, 0, :, :, : self.out_ch].permute(0, 3, 1, 2).contiguous()
