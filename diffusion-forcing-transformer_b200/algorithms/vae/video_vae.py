"""Decoder of the reference's causal VideoVAE on the B200 kernels — the step that follows the sampling path for latent
configurations.  Mirrors, on the decode side only,

  VideoVAE.__init__ / decode / _decode / from_pretrained   algorithms/vae/video_vae/model.py:282-342, 449-481, 505-530
  Decoder.forward                                          algorithms/vae/video_vae/model.py:252-270
  PaddedConv3D (causal, first frame repeated)              algorithms/vae/common/modules/conv.py:39-108
  ResnetBlock3D / AttnBlock3D                              algorithms/vae/common/modules/{resnet.py:93-109, attention.py:115-156}
  SpatialUpsample2x / Spatial2xTime2x3DUpsample            algorithms/vae/common/modules/updownsample.py:73-80, 131-147

with the same constructor arguments and the same state-dict keys (`decoder.*`, `post_quant_conv.*`), so a reference
checkpoint loads (encoder / quant_conv / loss keys are dropped: encoding is out of scope and raises).

Execution model (not a translation): activations are channel-last clips with a padded frame axis [B, 2 + T, H, W, C] —
an fp32 residual stream and bf16 conv operands whose two leading slots per clip hold copies of the clip's first frame,
so every causal 3x3x3 convolution of the whole batch is ONE implicit-GEMM launch over B*(2+T) frames (the temporal tap is
a frame offset of the TMA box; the two outputs that straddle a clip boundary land on the next clip's unused pad slots).
GroupNorm runs over the valid frames through a per-clip stride; the conv epilogues add bias and residual in fp32.
The mid-block attention (single head, d = C, per frame) is QK^T / softmax / PV on the tcgen05 GEMM with a row-softmax
kernel between them; the value bias is added after PV (softmax rows sum to one).  CUDA only — no CPU implementation.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from ... import ops
from ...checkpoint_io import load_checkpoint_file

PAD = 2          # leading pad slots per clip (the causal window of a kt = 3 convolution)
_CMIN = 64       # channel padding of the z / post-quant operands (one K tile of the implicit GEMM)


def _decoder_params(hidden_size: int, z_channels: int, embed_dim: int, mult: Tuple[int, ...], num_res_blocks: int,
                    use_quant_layer: bool) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) of the decode-side parameters in the reference's registration order (model.py:153-250)."""
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(name, cin, cout, k):
        out.extend([(f"{name}.conv.weight", (cout, cin) + k), (f"{name}.conv.bias", (cout,))])

    def norm(name, c):
        out.extend([(f"{name}.weight", (c,)), (f"{name}.bias", (c,))])

    def res(name, cin, cout):
        norm(f"{name}.norm1", cin)
        conv(f"{name}.conv1", cin, cout, (3, 3, 3))
        norm(f"{name}.norm2", cout)
        conv(f"{name}.conv2", cout, cout, (3, 3, 3))
        if cin != cout:
            conv(f"{name}.nin_shortcut", cin, cout, (1, 1, 1))

    L = len(mult)
    ch = hidden_size * mult[-1]
    conv("decoder.conv_in", z_channels, ch, (3, 3, 3))
    res("decoder.mid.block_1", ch, ch)
    norm("decoder.mid.attn_1.norm", ch)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", ch, ch, (1, 1, 1))
    res("decoder.mid.block_2", ch, ch)
    levels = {}
    for lvl in reversed(range(L)):
        cout, blocks = hidden_size * mult[lvl], []
        for i in range(num_res_blocks + 1):
            blocks.append((f"decoder.up.{lvl}.block.{i}", ch, cout))
            ch = cout
        levels[lvl] = (blocks, ch)
    for lvl in range(L):                                   # ModuleList order after up.insert(0, ...)
        blocks, c = levels[lvl]
        for b in blocks:
            res(*b)
        if lvl >= 1:
            conv(f"decoder.up.{lvl}.upsample.conv", c, c, (1, 3, 3) if lvl == 1 else (3, 3, 3))
    norm("decoder.norm_out", hidden_size * mult[0])
    conv("decoder.conv_out", hidden_size * mult[0], 3, (3, 3, 3))
    if use_quant_layer:
        conv("post_quant_conv", embed_dim, z_channels, (1, 1, 1))
    return out


def _register(root: nn.Module, key: str, p: nn.Parameter) -> None:
    *path, leaf = key.split(".")
    m = root
    for name in path:
        if name not in m._modules:
            m.add_module(name, nn.Module())
        m = m._modules[name]
    m.register_parameter(leaf, p)


class _DecoderOnKernels(nn.Module):
    """What VideoVAE and ImageVAE share: weights in kernel layout, the buffer pool and the building blocks on channel-last
    clips [B, pad + T, H, W, C] (pad = 2 for the causal 3-D decoder; pad = 0, T = 1 for plain image batches).
    `_conv_key(name)` maps a module name to its state-dict prefix (PaddedConv3D wraps its Conv3d in `.conv`)."""
    pad = PAD

    def _conv_key(self, name: str) -> str:
        return f"{name}.conv"

    def _init_runtime(self) -> None:
        self._packed: Optional[Dict] = None
        self._ws: Dict = {}

    def encode(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__}.encode is outside the scope of dfot_b200 (offline latents; "
                                  "decode only)")

    forward = encode

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        keep = {k: v for k, v in state_dict.items() if k.startswith(("decoder.", "post_quant_conv."))}
        self._packed = None
        return super().load_state_dict(keep, strict=strict, **kw)

    def _apply(self, fn, *a, **k):
        self._packed = None                                     # .to(device) / .float(): repack on the next decode
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------ weights in kernel layout
    def _pack(self, dev) -> Dict:
        if self._packed is not None and self._packed["dev"] == dev:
            return self._packed
        sd = {k: v.detach().to(dev, torch.float32) for k, v in self.state_dict().items()}
        P: Dict = {"dev": dev}

        def conv(name, cin_pad=None, cout_pad=None):
            w, b = sd[f"{self._conv_key(name)}.weight"], sd[f"{self._conv_key(name)}.bias"]
            if w.ndim == 4:
                w = w.unsqueeze(2)                                      # Conv2d: a kt = 1 temporal kernel
            cout, cin = w.shape[:2]
            ci, co = cin_pad or cin, cout_pad or cout
            wp = torch.zeros((co, ci) + tuple(w.shape[2:]), device=dev)
            wp[:cout, :cin] = w
            bp = torch.zeros((co,), device=dev)
            bp[:cout] = b
            if tuple(w.shape[2:]) == (1, 1, 1):
                wk = wp.reshape(co, ci)
            else:
                wk = wp.permute(0, 2, 3, 4, 1)                          # [Cout, kt, 3, 3, Cin]
            P[name] = (wk.contiguous().to(torch.bfloat16), bp.contiguous())

        def norm(name):
            P[name] = (sd[f"{name}.weight"].contiguous(), sd[f"{name}.bias"].contiguous())

        suffix = self._conv_key("") + ".weight"                         # ".conv.weight" (PaddedConv3D) or ".weight" (Conv2d)
        for key, w in sd.items():
            if not key.endswith(".weight"):
                continue
            if w.ndim == 1:
                norm(key[: -len(".weight")])
                continue
            name = key[: -len(suffix)]
            if name == "post_quant_conv":
                conv(name, _CMIN, _CMIN)
            elif name == "decoder.conv_in":
                conv(name, cin_pad=_CMIN)
            elif name == "decoder.conv_out":
                conv(name, cout_pad=8 * ((w.shape[0] + 7) // 8))
            else:
                conv(name)
        self._packed = P
        return P

    def _buf(self, tag: str, shape, dtype, dev) -> Tensor:
        key = (tag, tuple(shape), dtype, str(dev))
        t = self._ws.get(key)
        if t is None:
            t = self._ws[key] = torch.zeros(shape, dtype=dtype, device=dev)
        return t

    # ------------------------------------------------------------------ building blocks on padded clips
    def _valid(self, t: Tensor, frame: int) -> Tensor:
        return t.view(-1)[self.pad * frame:]

    def _gn(self, P, name, x: Tensor, out16: Tensor, B, T, HW, C, silu=True):
        """out16 valid frames <- [silu](GroupNorm(x valid frames)); pad slots <- first frame."""
        frame = HW * C
        sums = self._buf("gn", (B, 32, 3), torch.float64, x.device)
        xv = self._valid(x, frame)
        ops.groupnorm_stats_strided(xv, sums, B, T * HW, (self.pad + T) * frame, C)
        g, b = P[name]
        ops.groupnorm_apply_bf16(xv, sums, g, b, self._valid(out16, frame), B, T * HW, (self.pad + T) * frame, C, silu=silu)
        if self.pad:
            ops.vae_fill_pad_frames(out16, B, T, frame)

    def _conv(self, P, name, a16: Tensor, out: Tensor, B, T, H, W, resid: Optional[Tensor] = None):
        """out (fp32 clip) valid frames <- causal conv of the bf16 clip a16 (+ bias, + resid)."""
        w, bias = P[name]
        cin, cout, kt = w.shape[-1], w.shape[0], w.shape[1]
        n_all = B * (self.pad + T)
        x = a16.view(n_all, H, W, cin)
        if kt == 1:
            ops.conv3d_causal_bf16(x, w, out.view(-1), ops.EPI_F32 if resid is None else ops.EPI_RESID_F32, bias=bias,
                                   resid=None if resid is None else resid.view(-1))
            return
        frame = H * W * cout
        ops.conv3d_causal_bf16(x, w, self._valid(out, frame), ops.EPI_F32 if resid is None else ops.EPI_RESID_F32,
                               bias=bias, resid=None if resid is None else self._valid(resid, frame))

    def _resblock(self, P, name, x: Tensor, B, T, H, W, cin, cout) -> Tensor:
        dev, HW = x.device, H * W
        shape_in, shape_out = (B, self.pad + T, H, W, cin), (B, self.pad + T, H, W, cout)
        a16 = self._buf("a16", shape_in, torch.bfloat16, dev)
        self._gn(P, f"{name}.norm1", x, a16, B, T, HW, cin)
        h = self._buf("h", shape_out, torch.float32, dev)
        self._conv(P, f"{name}.conv1", a16, h, B, T, H, W)
        b16 = self._buf("a16", shape_out, torch.bfloat16, dev) if cin != cout else a16
        self._gn(P, f"{name}.norm2", h, b16, B, T, HW, cout)
        if cin != cout:                                            # nin_shortcut: 1x1x1 conv of the raw input
            x16 = self._buf("x16", shape_in, torch.bfloat16, dev)
            ops.cast_bf16(x, x16)
            sc = self._buf("sc", shape_out, torch.float32, dev)
            w, bias = P[f"{name}.nin_shortcut"]
            ops.gemm_bf16(x16.view(-1, cin), w, sc.view(-1, cout), ops.EPI_F32, bias=bias)
            x = sc
        y = self._buf("y0", shape_out, torch.float32, dev)
        if y.data_ptr() == x.data_ptr():
            y = self._buf("y1", shape_out, torch.float32, dev)
        self._conv(P, f"{name}.conv2", b16, y, B, T, H, W, resid=x)
        return y

    def _attn(self, P, name, x: Tensor, B, T, H, W, C) -> Tensor:
        dev, HW = x.device, H * W
        n_all = B * (self.pad + T)
        rows = n_all * HW
        if HW % 8 or HW > 1024:
            raise NotImplementedError("VideoVAE attention: H*W of the latent grid must be a multiple of 8 and <= 1024")
        a16 = self._buf("a16", (B, self.pad + T, H, W, C), torch.bfloat16, dev)
        self._gn(P, f"{name}.norm", x, a16, B, T, HW, C, silu=False)
        a = a16.view(rows, C)
        q = self._buf("q", (rows, C), torch.bfloat16, dev)
        k = self._buf("k", (rows, C), torch.bfloat16, dev)
        vt = self._buf("vt", (C, rows), torch.bfloat16, dev)
        ops.gemm_bf16(a, P[f"{name}.q"][0], q, ops.EPI_BF16, bias=P[f"{name}.q"][1])
        ops.gemm_bf16(a, P[f"{name}.k"][0], k, ops.EPI_BF16, bias=P[f"{name}.k"][1])
        ops.gemm_bf16(P[f"{name}.v"][0], a, vt, ops.EPI_BF16)      # V^T = W_v h^T for all frames at once, bias after PV
        s = self._buf("s", (rows, HW), torch.float32, dev)
        p = self._buf("p", (rows, HW), torch.bfloat16, dev)
        o = self._buf("o", (rows, C), torch.bfloat16, dev)
        frames = [b * (self.pad + T) + self.pad + t for b in range(B) for t in range(T)]
        for f in frames:
            r = slice(f * HW, (f + 1) * HW)
            ops.gemm_bf16(q[r], k[r], s[r], ops.EPI_F32)
        ops.softmax_rows_bf16(s, p, scale=float(int(C) ** -0.5))
        vb = P[f"{name}.v"][1]
        for f in frames:
            r = slice(f * HW, (f + 1) * HW)
            ops.gemm_bf16(p[r], vt[:, r], o[r], ops.EPI_BF16, bias=vb)
        y = self._buf("y0", (B, self.pad + T, H, W, C), torch.float32, dev)
        if y.data_ptr() == x.data_ptr():
            y = self._buf("y1", (B, self.pad + T, H, W, C), torch.float32, dev)
        w, bias = P[f"{name}.proj_out"]
        ops.gemm_bf16(o, w, y.view(rows, C), ops.EPI_RESID_F32, bias=bias, resid=x.view(rows, C))
        return y

    # ------------------------------------------------------------------ the decoder walk (Decoder.forward of both VAEs)
    def _run_decoder(self, z_cl: Tensor, B, T, H, W, Cz, quant: bool = True):
        """z_cl [B, T, H, W, Cz] (any strides) -> (fp32 clip [B, pad + T', H', W', 8k] whose first channels are the
        decoded ones, T')."""
        dev, pad = z_cl.device, self.pad
        P = self._pack(dev)
        sig = (B, T, H, W, str(dev))
        if self._ws.get("sig") != sig:                          # one workspace set per input shape: a new shape frees the old
            self._ws = {"sig": sig}
        # latent -> channel-last (padded) bf16 clip (layout change of a tiny tensor; channels zero-padded to one K tile)
        z16 = self._buf("z16", (B, pad + T, H, W, _CMIN), torch.bfloat16, dev)
        z16[:, pad:, :, :, :Cz] = z_cl
        if pad:
            ops.vae_fill_pad_frames(z16, B, T, H * W * _CMIN)
        if quant:
            w, bias = P["post_quant_conv"]
            zq = self._buf("zq16", (B, pad + T, H, W, _CMIN), torch.bfloat16, dev)
            ops.gemm_bf16(z16.view(-1, _CMIN), w, zq.view(-1, _CMIN), ops.EPI_BF16, bias=bias)
            z16 = zq
        ch = self.hidden_size * self.mult[-1]
        h = self._buf("y0", (B, pad + T, H, W, ch), torch.float32, dev)
        self._conv(P, "decoder.conv_in", z16, h, B, T, H, W)
        h = self._resblock(P, "decoder.mid.block_1", h, B, T, H, W, ch, ch)
        h = self._attn(P, "decoder.mid.attn_1", h, B, T, H, W, ch)
        h = self._resblock(P, "decoder.mid.block_2", h, B, T, H, W, ch, ch)
        for lvl in reversed(range(len(self.mult))):
            cout = self.hidden_size * self.mult[lvl]
            for i in range(self.nrb + 1):
                h = self._resblock(P, f"decoder.up.{lvl}.block.{i}", h, B, T, H, W, ch, cout)
                ch = cout
            up = self._upsample(lvl, h, B, T, H, W, ch)
            if up is not None:
                u16, T = up
                H, W = 2 * H, 2 * W
                h = self._buf("y0", (B, pad + T, H, W, ch), torch.float32, dev)
                self._conv(P, f"decoder.up.{lvl}.upsample.conv", u16, h, B, T, H, W)
        a16 = self._buf("a16", (B, pad + T, H, W, ch), torch.bfloat16, dev)
        self._gn(P, "decoder.norm_out", h, a16, B, T, H * W, ch)
        cpad = P["decoder.conv_out"][0].shape[0]
        out = self._buf("out", (B, pad + T, H, W, cpad), torch.float32, dev)
        self._conv(P, "decoder.conv_out", a16, out, B, T, H, W)
        return out, T


class VideoVAE(_DecoderOnKernels):
    """Decode side of the reference's VideoVAE (default topology: 3-D ResNet blocks everywhere, attention in the mid
    block only, level 1 upsampled spatially, levels 2.. spatially and temporally)."""

    def __init__(self, hidden_size: int = 128, z_channels: int = 4, hidden_size_mult: Tuple[int, ...] = (1, 2, 4, 4),
                 attn_resolutions: Tuple[int, ...] = (), dropout: float = 0.0, resolution: int = 256,
                 temporal_length: int = 17, double_z: bool = True, embed_dim: int = 4, num_res_blocks: int = 2,
                 use_quant_layer: bool = True, is_causal: bool = True, first_padding_mode: str = "same", **topology):
        super().__init__()
        defaults = dict(q_conv="PaddedConv3D", decoder_conv_in="PaddedConv3D", decoder_conv_out="PaddedConv3D",
                        decoder_attention="AttnBlock3D", decoder_resnet_blocks=("ResnetBlock3D",) * 4,
                        decoder_spatial_upsample=("", "SpatialUpsample2x", "Spatial2xTime2x3DUpsample",
                                                  "Spatial2xTime2x3DUpsample"),
                        decoder_temporal_upsample=("",) * 4, decoder_mid_resnet="ResnetBlock3D")
        for k, v in topology.items():
            if k.startswith("encoder_"):
                continue                                    # the encoder is never built here
            if k not in defaults or (tuple(v) if isinstance(v, (list, tuple)) else v) != defaults[k]:
                raise NotImplementedError(f"VideoVAE: decoder topology {k}={v!r} is not the reference default")
        if not is_causal or first_padding_mode != "same" or tuple(attn_resolutions) or len(hidden_size_mult) != 4:
            raise NotImplementedError("VideoVAE: only the causal, first-frame-padded, 4-level default decoder is built")
        if hidden_size % 32:
            raise ValueError("VideoVAE: hidden_size must be a multiple of 32 (GroupNorm groups)")
        self.hidden_size, self.z_channels, self.embed_dim = hidden_size, z_channels, embed_dim
        self.mult, self.nrb = tuple(hidden_size_mult), num_res_blocks
        self.use_quant_layer, self.is_causal, self.temporal_length = use_quant_layer, is_causal, temporal_length
        self.temporal_latent_length = (temporal_length - 1) // 4 + 1
        g = torch.Generator().manual_seed(0)
        for key, shape in _decoder_params(hidden_size, z_channels, embed_dim, self.mult, num_res_blocks, use_quant_layer):
            if key.endswith("conv.weight"):
                bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3] * shape[4])
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            elif key.endswith("conv.bias"):
                t = torch.zeros(shape)
            else:
                t = torch.ones(shape) if key.endswith(".weight") else torch.zeros(shape)
            _register(self, key, nn.Parameter(t, requires_grad=False))
        self._init_runtime()

    # ------------------------------------------------------------------ checkpoint (model.py:505-530)
    @classmethod
    def from_pretrained(cls, path: str, **kwargs) -> "VideoVAE":
        ckpt = load_checkpoint_file(path)
        cfg = {k: tuple(v) if isinstance(v, list) else v for k, v in ckpt["model_cfg"].items()}
        model = cls(**cfg)
        own = [n for n, _ in model.named_parameters()]
        if len(ckpt.get("optimizer_states", [])) > 0 and "ema" in ckpt["optimizer_states"][0]:
            # EMA weights are stored as a list in the FULL model's named_parameters() order; names are needed to pick the
            # decode side, so the checkpoint's own state_dict supplies them
            names = [k.replace("vae.", "", 1) for k in ckpt["state_dict"] if k.startswith("vae.")]
            ema = ckpt["optimizer_states"][0]["ema"]
            if len(names) != len(ema):
                raise RuntimeError(f"VideoVAE.from_pretrained: {len(ema)} EMA tensors for {len(names)} `vae.*` entries of "
                                   "the state_dict — cannot name the EMA weights")
            full = dict(zip(names, ema))
        else:
            full = {k.replace("vae.", "", 1): v for k, v in ckpt["state_dict"].items() if k.startswith("vae.")}
        missing = [n for n in own if n not in full]
        if missing:
            raise RuntimeError(f"VideoVAE.from_pretrained: checkpoint lacks decoder tensors {missing[:4]} ...")
        model.load_state_dict({n: full[n] for n in own})
        return model

    # ------------------------------------------------------------------ decode (model.py:449-481, 252-270)
    def _upsample(self, lvl: int, h: Tensor, B, T, H, W, ch):
        if lvl < 1:
            return None
        temporal = lvl >= 2                                        # SpatialUpsample2x at level 1, Spatial2xTime2x above
        To = 2 * T - 1 if temporal else T
        u16 = self._buf("a16", (B, PAD + To, 2 * H, 2 * W, ch), torch.bfloat16, h.device)
        ops.vae_upsample2x_bf16(h, u16, B, T, H, W, ch, temporal)
        return u16, To

    @torch.no_grad()
    def decode(self, z: Tensor, desired_length: Optional[int] = None) -> Tensor:
        """z [B, C_z, T, H, W] -> video [B, 3, 1 + 4 (T - 1), 8 H, 8 W] fp32 (the last `desired_length` frames)."""
        ops.require_cuda(z.device, "VideoVAE.decode")
        B, Cz, T, H, W = z.shape
        if Cz != (self.embed_dim if self.use_quant_layer else self.z_channels):
            raise ValueError(f"VideoVAE.decode: expected {self.embed_dim} latent channels, got {Cz}")
        out, T = self._run_decoder(z.permute(0, 2, 3, 4, 1), B, T, H, W, Cz, self.use_quant_layer)
        video = out[:, PAD:, :, :, :3].permute(0, 4, 1, 2, 3).contiguous()
        if desired_length is not None:
            video = video[:, :, -desired_length:]
            assert video.shape[2] == desired_length, \
                f"Desired length {desired_length} does not match decoded length {video.shape[2]}"
        return video
