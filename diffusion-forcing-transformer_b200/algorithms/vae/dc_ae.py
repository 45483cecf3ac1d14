"""Decoder of the DC-AE image autoencoder on the B200 kernels — the VAE of the DMLab / Minecraft latent configurations
(`vae.name: dc_ae_preprocessor`, selected at algorithms/common/base_pytorch_video_algo.py:511-520).
Mirrors, on the decode side only, the reference's algorithms/vae/dc_ae/autoencoder_dc_model.py

    MyAutoencoderDC.__init__ / from_pretrained / decode      :474-553, 705-737, 643-665
    Decoder.forward                                          :453-468
    DCUpBlock2d / ResBlock / EfficientViTBlock               :222-260 / :109-136 / :139-172
    SanaMultiscaleLinearAttention.apply_linear_attention     :87-96

and the diffusers==0.32.2 modules that file imports (GLUMBConv, RMSNorm, SanaMultiscaleAttnProcessor2_0), with the same
`cfg` tree (configurations/algorithm/dc_ae_preprocessor.yaml) and the same state-dict keys (`decoder.*`, BatchNorm buffers
included), so a DC-AE checkpoint loads unchanged.  Execution model (channel-last [n, H, W, C] everywhere):

  * every 3x3 convolution is the implicit-GEMM tcgen05 kernel, every 1x1 convolution / linear the tcgen05 GEMM;
  * eval-mode BatchNorm2d after a ResBlock's bias-free conv2 is an affine per output channel: folded into conv2's weights
    and bias when the weights are packed, so a ResBlock is conv(+bias) -> ReLU -> conv with the residual epilogue;
  * pixel-shuffle + channel-repeat shortcut, ReLU linear attention, depthwise 3x3 + GLU and RMSNorm(+residual / ReLU)
    are the small kernels of csrc/dcae.cu.
The encoder half is not built (latents come from the dataset).  CUDA only.
"""
import copy
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from ... import ops
from ...checkpoint_io import load_checkpoint_file
from ...config import to_config
from .video_vae import _register


def _pad(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def decoder_layout(cfg) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) of the decoder's state dict in the reference's registration order; kind in
    {"param", "buffer", "count"} (BatchNorm running statistics / num_batches_tracked are buffers)."""
    ch, layers = list(cfg.decoder_block_out_channels), list(cfg.decoder_layers_per_block)
    n, lat = len(ch), cfg.latent_channels
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    P = lambda k, s: out.append((k, tuple(s), "param"))

    def norm(pre, c, kind):
        P(pre + ".weight", (c,))
        P(pre + ".bias", (c,))
        if kind != "rms_norm":
            out.append((pre + ".running_mean", (c,), "buffer"))
            out.append((pre + ".running_var", (c,), "buffer"))
            out.append((pre + ".num_batches_tracked", (), "count"))

    P("decoder.conv_in.weight", (ch[-1], lat, 3, 3))
    P("decoder.conv_in.bias", (ch[-1],))
    for i in range(n):
        k, c = 0, ch[i]
        if i < n - 1 and layers[i] > 0:
            P(f"decoder.up_blocks.{i}.{k}.conv.weight", (4 * c, ch[i + 1], 3, 3))
            P(f"decoder.up_blocks.{i}.{k}.conv.bias", (4 * c,))
            k += 1
        for _ in range(layers[i]):
            pre = f"decoder.up_blocks.{i}.{k}"
            if cfg.decoder_block_types[i] == "ResBlock":
                P(pre + ".conv1.weight", (c, c, 3, 3))
                P(pre + ".conv1.bias", (c,))
                P(pre + ".conv2.weight", (c, c, 3, 3))
                norm(pre + ".norm", c, cfg.decoder_norm_types[i])
            else:
                for nm in ("to_q", "to_k", "to_v", "to_out"):
                    P(pre + f".attn.{nm}.weight", (c, c))
                norm(pre + ".attn.norm_out", c, cfg.decoder_norm_types[i])
                g = pre + ".conv_out"
                P(g + ".conv_inverted.weight", (8 * c, c, 1, 1))
                P(g + ".conv_inverted.bias", (8 * c,))
                P(g + ".conv_depth.weight", (8 * c, 1, 3, 3))
                P(g + ".conv_depth.bias", (8 * c,))
                P(g + ".conv_point.weight", (c, 4 * c, 1, 1))
                norm(g + ".norm", c, "rms_norm")
            k += 1
    c0 = ch[0] if layers[0] > 0 else ch[1]
    norm("decoder.norm_out", c0, "rms_norm")
    if layers[0] > 0:
        P("decoder.conv_out.weight", (cfg.in_channels, c0, 3, 3))
        P("decoder.conv_out.bias", (cfg.in_channels,))
    else:
        P("decoder.conv_out.conv.weight", (4 * cfg.in_channels, c0, 3, 3))
        P("decoder.conv_out.conv.bias", (4 * cfg.in_channels,))
    return out


def _register_buffer(root: nn.Module, key: str, t: Tensor) -> None:
    *path, leaf = key.split(".")
    m = root
    for name in path:
        if name not in m._modules:
            m.add_module(name, nn.Module())
        m = m._modules[name]
    m.register_buffer(leaf, t)


class MyAutoencoderDC(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        cfg = copy.deepcopy(to_config(cfg))
        self.cfg = cfg
        n = len(cfg.decoder_block_out_channels)
        as_list = lambda v: list(v) if isinstance(v, (list, tuple)) else [v] * n
        cfg.decoder_block_types, cfg.decoder_norm_types, cfg.decoder_act_fns = as_list(cfg.decoder_block_types), \
            as_list(cfg.decoder_norm_types), as_list(cfg.decoder_act_fns)
        bad = []
        if any(len(m) for m in cfg.decoder_qkv_multiscales):
            bad.append("decoder_qkv_multiscales (multi-scale depthwise projections)")
        if cfg.get("upsample_block_type", "pixel_shuffle") != "pixel_shuffle":
            bad.append("upsample_block_type=interpolate")
        if any(t not in ("ResBlock", "EfficientViTBlock") for t in cfg.decoder_block_types):
            bad.append("block types other than ResBlock / EfficientViTBlock")
        if any(a not in ("relu", "silu") for a, t in zip(cfg.decoder_act_fns, cfg.decoder_block_types) if t == "ResBlock"):
            bad.append("ResBlock activations other than relu / silu")
        if cfg.attention_head_dim > 32:
            bad.append("attention_head_dim > 32")
        if bad:
            raise NotImplementedError(f"MyAutoencoderDC: not built by dfot_b200: {bad}")
        self.scaling_factor = cfg.scaling_factor
        self.spatial_compression_ratio = 2 ** (n - 1)
        self.temporal_compression_ratio = 1
        g = torch.Generator().manual_seed(0)
        for key, shape, kind in decoder_layout(cfg):
            if kind == "count":
                _register_buffer(self, key, torch.tensor(0, dtype=torch.int64))
            elif kind == "buffer":
                _register_buffer(self, key, torch.ones(shape) if key.endswith("running_var") else torch.zeros(shape))
            else:
                if len(shape) >= 2:
                    fan_in = 1
                    for s in shape[1:]:
                        fan_in *= s
                    t = (torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5
                else:
                    t = torch.ones(shape) if (".norm" in key and key.endswith(".weight")) else torch.zeros(shape)
                _register(self, key, nn.Parameter(t, requires_grad=False))
        self._packed: Optional[Dict] = None
        self._ws: Dict = {}

    # ------------------------------------------------------------------ checkpoint (:705-737)
    @classmethod
    def from_pretrained(cls, cfg, **kwargs) -> "MyAutoencoderDC":
        cfg = to_config(cfg)
        path = cfg.pretrained_path
        if path.startswith("diffuser:"):
            raise NotImplementedError("MyAutoencoderDC: diffusers hub checkpoints need network access; pass a local file")
        if path.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd = load_file(path)
        else:
            sd = load_checkpoint_file(path)
        model = cls(cfg)
        model.load_state_dict(sd)          # decode side strict, encoder tensors ignored (the reference: strict=False)
        return model.eval()

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        keep = {k: v for k, v in state_dict.items() if k.startswith("decoder.")}
        self._packed = None
        return super().load_state_dict(keep, strict=strict, **kw)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def encode(self, *a, **k):
        raise NotImplementedError("MyAutoencoderDC.encode is outside the scope of dfot_b200 (offline latents; decode only)")

    forward = encode

    # ------------------------------------------------------------------ weights in kernel layout
    def _pack(self, dev) -> Dict:
        if self._packed is not None and self._packed["dev"] == dev:
            return self._packed
        sd = {k: v.detach().to(dev, torch.float32) for k, v in self.state_dict().items() if v.is_floating_point()}
        cfg = self.cfg
        P: Dict = {"dev": dev}
        bf = lambda w: w.contiguous().to(torch.bfloat16)

        def conv3(w, b, cin_pad=None, cout_pad=None):
            """[Cout, Cin, 3, 3] (+ bias) -> implicit-GEMM operand [Cout', 3, 3, Cin'] bf16, bias f32 [Cout']."""
            cout, cin = w.shape[:2]
            ci, co = cin_pad or _pad(cin, 8), cout_pad or _pad(cout, 8)
            wp = torch.zeros((co, 3, 3, ci), device=dev)
            wp[:cout, :, :, :cin] = w.permute(0, 2, 3, 1)
            bp = torch.zeros((co,), device=dev)
            if b is not None:
                bp[:cout] = b
            return bf(wp), bp.contiguous()

        P["conv_in"] = conv3(sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"])
        ch, layers = list(cfg.decoder_block_out_channels), list(cfg.decoder_layers_per_block)
        n = len(ch)
        stages = []
        for i in reversed(range(n)):
            k, c, blocks = 0, ch[i], []
            up = None
            if i < n - 1 and layers[i] > 0:
                pre = f"decoder.up_blocks.{i}.{k}.conv"
                up = conv3(sd[pre + ".weight"], sd[pre + ".bias"]) + (4 * c // ch[i + 1],)       # (w, b, repeats)
                k += 1
            for _ in range(layers[i]):
                pre = f"decoder.up_blocks.{i}.{k}"
                if cfg.decoder_block_types[i] == "ResBlock":
                    w2 = sd[pre + ".conv2.weight"]
                    d = dict(kind="res", act=cfg.decoder_act_fns[i], conv1=conv3(sd[pre + ".conv1.weight"], sd[pre + ".conv1.bias"]))
                    if cfg.decoder_norm_types[i] == "rms_norm":
                        d.update(conv2=conv3(w2, None), norm=(sd[pre + ".norm.weight"].contiguous(), sd[pre + ".norm.bias"].contiguous()))
                    else:   # BatchNorm2d (eval, eps 1e-5): y = (conv - mean) * gamma / sqrt(var + eps) + beta — folded
                        s = sd[pre + ".norm.weight"] * torch.rsqrt(sd[pre + ".norm.running_var"] + 1e-5)
                        d.update(conv2=conv3(w2 * s[:, None, None, None], sd[pre + ".norm.bias"] - sd[pre + ".norm.running_mean"] * s),
                                 norm=None)
                else:
                    a, g = pre + ".attn", pre + ".conv_out"
                    d = dict(kind="vit", qkv_w=bf(torch.cat([sd[a + f".to_{x}.weight"] for x in "qkv"], 0)),
                             out_w=bf(sd[a + ".to_out.weight"]),
                             attn_norm=(sd[a + ".norm_out.weight"].contiguous(), sd[a + ".norm_out.bias"].contiguous()),
                             inv_w=bf(sd[g + ".conv_inverted.weight"].reshape(8 * c, c)), inv_b=sd[g + ".conv_inverted.bias"].contiguous(),
                             dw_w=sd[g + ".conv_depth.weight"].reshape(8 * c, 9).contiguous(), dw_b=sd[g + ".conv_depth.bias"].contiguous(),
                             pt_w=bf(sd[g + ".conv_point.weight"].reshape(c, 4 * c)),
                             glu_norm=(sd[g + ".norm.weight"].contiguous(), sd[g + ".norm.bias"].contiguous()))
                    if cfg.decoder_norm_types[i] != "rms_norm":
                        raise NotImplementedError("MyAutoencoderDC: EfficientViTBlock with batch_norm is not built")
                blocks.append(d)
                k += 1
            stages.append(dict(c=c, up=up, blocks=blocks))
        P["stages"] = stages
        P["norm_out"] = (sd["decoder.norm_out.weight"].contiguous(), sd["decoder.norm_out.bias"].contiguous())
        if layers[0] > 0:
            P["conv_out"], P["out_shuffle"] = conv3(sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"]), False
        else:
            P["conv_out"], P["out_shuffle"] = conv3(sd["decoder.conv_out.conv.weight"], sd["decoder.conv_out.conv.bias"]), True
        self._packed = P
        return P

    def _buf(self, name: str, shape, dtype, dev) -> Tensor:
        key = (name, tuple(shape), dtype, str(dev))
        t = self._ws.get(key)
        if t is None:
            t = self._ws[key] = torch.empty(shape, dtype=dtype, device=dev)
        return t

    # ------------------------------------------------------------------ decode (:643-665, 453-468)
    @torch.no_grad()
    def decode(self, z: Tensor) -> Tensor:
        """z [n, latent_channels, H, W] -> images [n, in_channels, 2^(L-1) H, 2^(L-1) W] fp32."""
        ops.require_cuda(z.device, "MyAutoencoderDC.decode")
        cfg, dev = self.cfg, z.device
        n, Cz, H, W = z.shape
        if Cz != cfg.latent_channels:
            raise ValueError(f"MyAutoencoderDC.decode: expected {cfg.latent_channels} latent channels, got {Cz}")
        P = self._pack(dev)
        f32, b16 = torch.float32, torch.bfloat16
        ch = list(cfg.decoder_block_out_channels)
        # conv_in + in_shortcut: z.repeat_interleave(C / latent) is the residual operand of the convolution
        zc = z.float().permute(0, 2, 3, 1).contiguous()                                  # [n, H, W, Cz]
        w_in, b_in = P["conv_in"]
        z16 = torch.zeros((n, H, W, w_in.shape[-1]), dtype=b16, device=dev)
        z16[..., :Cz] = zc.to(b16)
        C = ch[-1]
        x = self._buf("x", (n * H * W, C), f32, dev)
        ops.conv3x3_bf16(z16, w_in, x, ops.EPI_RESID_F32, bias=b_in,
                         resid=zc.repeat_interleave(C // Cz, dim=-1).reshape(n * H * W, C).contiguous())
        for st in P["stages"]:
            c = st["c"]
            if st["up"] is not None:
                w, b, repeats = st["up"]
                x16 = ops.cast_bf16(x, self._buf("x16", x.shape, b16, dev))
                y = self._buf("up", (n * H * W, w.shape[0]), f32, dev)
                ops.conv3x3_bf16(x16.view(n, H, W, C), w, y, ops.EPI_F32, bias=b)
                nx = self._buf("x", (n * 4 * H * W, c), f32, dev)
                ops.pixel_shuffle2x(y, c, n, H, W, shortcut=x, repeats=repeats, out_f32=nx)
                x, H, W, C = nx, 2 * H, 2 * W, c
            M = n * H * W
            for blk in st["blocks"]:
                if blk["kind"] == "res":
                    x16 = ops.cast_bf16(x, self._buf("x16", x.shape, b16, dev))
                    h16 = self._buf("h16", (M, C), b16, dev)
                    w1, b1 = blk["conv1"]
                    if blk["act"] == "silu":
                        ops.conv3x3_bf16(x16.view(n, H, W, C), w1, h16, ops.EPI_SILU_BF16, bias=b1)
                    else:
                        ops.conv3x3_bf16(x16.view(n, H, W, C), w1, h16, ops.EPI_BF16, bias=b1)
                        ops.relu_bf16(h16)
                    w2, b2 = blk["conv2"]
                    if blk["norm"] is None:         # BatchNorm folded: conv2 writes x + bn(conv2(h)) in place
                        ops.conv3x3_bf16(h16.view(n, H, W, C), w2, x, ops.EPI_RESID_F32, bias=b2, resid=x)
                    else:
                        t = self._buf("t", (M, C), f32, dev)
                        ops.conv3x3_bf16(h16.view(n, H, W, C), w2, t, ops.EPI_F32, bias=b2)
                        ops.rmsnorm_affine(t, *blk["norm"], 1e-5, resid=x, out_f32=x)
                else:
                    d, heads = cfg.attention_head_dim, C // cfg.attention_head_dim
                    x16 = ops.cast_bf16(x, self._buf("x16", x.shape, b16, dev))
                    qkv = self._buf("qkv", (M, 3 * C), f32, dev)
                    ops.gemm_bf16(x16, blk["qkv_w"], qkv, ops.EPI_F32)
                    att = self._buf("att", (M, C), f32, dev)
                    ops.linear_attention_relu(qkv, att, n, H * W, heads, d, 1e-15)
                    a16 = ops.cast_bf16(att, self._buf("a16", att.shape, b16, dev))
                    t = self._buf("t", (M, C), f32, dev)
                    ops.gemm_bf16(a16, blk["out_w"], t, ops.EPI_F32)
                    ops.rmsnorm_affine(t, *blk["attn_norm"], 1e-5, resid=x, out_f32=x)
                    # GLUMBConv: 1x1 expand (+SiLU) -> depthwise 3x3 + GLU -> 1x1 project -> RMSNorm + residual
                    x16 = ops.cast_bf16(x, self._buf("x16", x.shape, b16, dev))
                    e16 = self._buf("e16", (M, 8 * C), b16, dev)
                    ops.gemm_bf16(x16, blk["inv_w"], e16, ops.EPI_SILU_BF16, bias=blk["inv_b"])
                    g16 = self._buf("g16", (M, 4 * C), b16, dev)
                    ops.dwconv3x3_glu_bf16(e16, blk["dw_w"], blk["dw_b"], g16, n, H, W, 4 * C)
                    ops.gemm_bf16(g16, blk["pt_w"], t, ops.EPI_F32)
                    ops.rmsnorm_affine(t, *blk["glu_norm"], 1e-5, resid=x, out_f32=x)
        # norm_out (RMSNorm over channels) -> ReLU -> conv_out
        M = n * H * W
        y16 = self._buf("y16", (M, C), b16, dev)
        ops.rmsnorm_affine(x, *P["norm_out"], 1e-5, relu=True, out_bf16=y16)
        w, b = P["conv_out"]
        o = self._buf("o", (M, w.shape[0]), f32, dev)
        ops.conv3x3_bf16(y16.view(n, H, W, C), w, o, ops.EPI_F32, bias=b)
        Co = cfg.in_channels
        if P["out_shuffle"]:
            img = torch.empty((n, 2 * H, 2 * W, Co), dtype=f32, device=dev)
            ops.pixel_shuffle2x(o, Co, n, H, W, out_f32=img)
            return img.permute(0, 3, 1, 2).contiguous()
        return o.view(n, H, W, -1)[..., :Co].permute(0, 3, 1, 2).contiguous()
