"""ORACLE (test infrastructure): the DC-AE image decoder (the VAE of the DMLab / Minecraft latent configurations),
functional over a reference-keyed state dict, torch fp32 on the CPU.

Restates algorithms/vae/dc_ae/autoencoder_dc_model.py
  Decoder.forward :453-468 (conv_in + channel-repeat shortcut, up blocks in reverse, RMSNorm -> ReLU -> conv_out),
  DCUpBlock2d :222-260 (conv -> pixel_shuffle, shortcut = pixel_shuffle(repeat_interleave(x))),
  ResBlock :109-136 (conv1 -> act -> conv2 -> BatchNorm2d (eval) / RMSNorm -> + residual),
  EfficientViTBlock :139-172, SanaMultiscaleLinearAttention.apply_linear_attention :87-96,
and the third-party modules that file imports from diffusers==0.32.2 (requirements.txt:4), which is NOT installed here:
  SanaMultiscaleAttnProcessor2_0 (attention_processor.py), GLUMBConv (transformers/sana_transformer.py),
  RMSNorm / get_normalization (normalization.py), get_activation (activations.py).
Their published behaviour is restated at the reference's call sites (lines above); **parity for those four pieces is
unpinned** — oracle/make_goldens_dcae.py executes the reference's own Decoder / blocks, but through stand-ins for the
diffusers imports that carry the same restatement.
"""
from typing import Dict, Sequence

import torch
import torch.nn.functional as F


def rms_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """diffusers RMSNorm(dim, eps, elementwise_affine=True, bias=True) over the LAST dimension."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    return x * torch.rsqrt(var + eps) * w + b


def _act(name: str):
    return {"relu": F.relu, "silu": F.silu, "relu6": F.relu6, "gelu": F.gelu}[name]


class DCAEDecoderOracle:
    def __init__(self, sd: Dict[str, torch.Tensor], cfg: dict):
        self.sd = {k: v.detach().float() for k, v in sd.items()}
        self.latent = cfg["latent_channels"]
        self.channels = list(cfg["decoder_block_out_channels"])
        self.layers = list(cfg["decoder_layers_per_block"])
        n = len(self.channels)
        as_list = lambda v: list(v) if isinstance(v, (list, tuple)) else [v] * n
        self.types, self.norms, self.acts = as_list(cfg["decoder_block_types"]), as_list(cfg["decoder_norm_types"]), \
            as_list(cfg["decoder_act_fns"])
        self.head_dim = cfg["attention_head_dim"]
        assert all(len(m) == 0 for m in cfg["decoder_qkv_multiscales"]), "multi-scale qkv projections are not restated"
        assert cfg.get("upsample_block_type", "pixel_shuffle") == "pixel_shuffle"

    # ---- blocks
    def up_block(self, x, pre, shortcut=True):
        sd = self.sd
        y = F.pixel_shuffle(F.conv2d(x, sd[pre + ".conv.weight"], sd[pre + ".conv.bias"], padding=1), 2)
        if shortcut:
            repeats = sd[pre + ".conv.weight"].shape[0] // x.shape[1]        # out_channels * 4 / in_channels
            y = y + F.pixel_shuffle(x.repeat_interleave(repeats, dim=1), 2)
        return y

    def res_block(self, x, pre, norm, act):
        sd = self.sd
        h = F.conv2d(x, sd[pre + ".conv1.weight"], sd[pre + ".conv1.bias"], padding=1)
        h = F.conv2d(_act(act)(h), sd[pre + ".conv2.weight"], None, padding=1)
        if norm == "rms_norm":
            h = rms_norm(h.movedim(1, -1), sd[pre + ".norm.weight"], sd[pre + ".norm.bias"]).movedim(-1, 1)
        else:       # BatchNorm2d in eval mode (get_normalization("batch_norm"): eps 1e-5, affine)
            h = F.batch_norm(h, sd[pre + ".norm.running_mean"], sd[pre + ".norm.running_var"], sd[pre + ".norm.weight"],
                             sd[pre + ".norm.bias"], training=False, eps=1e-5)
        return h + x

    def vit_block(self, x, pre, norm):
        sd = self.sd
        B, C, H, W = x.shape
        d = self.head_dim
        # --- SanaMultiscaleAttnProcessor2_0: [q | k | v] concatenated, then read as groups of 3*d channels
        t = x.movedim(1, -1)
        qkv = torch.cat([F.linear(t, sd[pre + f".attn.to_{n}.weight"]) for n in "qkv"], dim=3).movedim(-1, 1)
        qkv = qkv.float().reshape(B, -1, 3 * d, H * W)
        q, k, v = qkv.chunk(3, dim=2)
        q, k = F.relu(q), F.relu(k)
        assert H * W > d, "quadratic attention (H*W <= head_dim) is not restated"
        v = F.pad(v, (0, 0, 0, 1), mode="constant", value=1)
        h = torch.matmul(torch.matmul(v, k.transpose(-1, -2)), q)
        h = h[:, :, :-1] / (h[:, :, -1:] + 1e-15)
        h = h.reshape(B, -1, H, W)
        h = F.linear(h.movedim(1, -1), sd[pre + ".attn.to_out.weight"])
        if norm == "rms_norm":
            h = rms_norm(h, sd[pre + ".attn.norm_out.weight"], sd[pre + ".attn.norm_out.bias"]).movedim(-1, 1)
        else:
            h = F.batch_norm(h.movedim(-1, 1), sd[pre + ".attn.norm_out.running_mean"], sd[pre + ".attn.norm_out.running_var"],
                             sd[pre + ".attn.norm_out.weight"], sd[pre + ".attn.norm_out.bias"], training=False, eps=1e-5)
        x = h + x
        # --- GLUMBConv (norm_type rms_norm, residual)
        g = pre + ".conv_out"
        h = F.silu(F.conv2d(x, sd[g + ".conv_inverted.weight"], sd[g + ".conv_inverted.bias"]))
        h = F.conv2d(h, sd[g + ".conv_depth.weight"], sd[g + ".conv_depth.bias"], padding=1, groups=h.shape[1])
        h, gate = torch.chunk(h, 2, dim=1)
        h = F.conv2d(h * F.silu(gate), sd[g + ".conv_point.weight"])
        h = rms_norm(h.movedim(1, -1), sd[g + ".norm.weight"], sd[g + ".norm.bias"]).movedim(-1, 1)
        return h + x

    # ---- Decoder.forward
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        sd = self.sd
        n = len(self.channels)
        x = F.conv2d(z.float(), sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1) + \
            z.float().repeat_interleave(self.channels[-1] // self.latent, dim=1)
        for i in reversed(range(n)):
            k = 0
            if i < n - 1 and self.layers[i] > 0:
                x = self.up_block(x, f"decoder.up_blocks.{i}.{k}")
                k += 1
            for _ in range(self.layers[i]):
                pre = f"decoder.up_blocks.{i}.{k}"
                x = self.res_block(x, pre, self.norms[i], self.acts[i]) if self.types[i] == "ResBlock" else \
                    self.vit_block(x, pre, self.norms[i])
                k += 1
        x = F.relu(rms_norm(x.movedim(1, -1), sd["decoder.norm_out.weight"], sd["decoder.norm_out.bias"]).movedim(-1, 1))
        if self.layers[0] > 0:
            return F.conv2d(x, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
        return self.up_block(x, "decoder.conv_out", shortcut=False)


def small_cfg() -> dict:
    """A DC-AE decoder with the topology of configurations/algorithm/dc_ae_preprocessor.yaml at test size."""
    return dict(in_channels=3, latent_channels=8, attention_head_dim=32, scaling_factor=0.2889,
                decoder_block_types=["ResBlock", "ResBlock", "ResBlock", "EfficientViTBlock"],
                decoder_block_out_channels=[32, 64, 128, 128], decoder_layers_per_block=[0, 2, 3, 2],
                decoder_norm_types=["batch_norm", "batch_norm", "batch_norm", "rms_norm"],
                decoder_act_fns=["relu", "relu", "relu", "silu"], decoder_qkv_multiscales=[[], [], [], []],
                upsample_block_type="pixel_shuffle",
                # (the encoder half exists in the reference's constructor; the decode path never touches it)
                encoder_block_types=["ResBlock", "ResBlock", "ResBlock", "EfficientViTBlock"],
                encoder_block_out_channels=[32, 64, 128, 128], encoder_layers_per_block=[0, 1, 1, 1],
                encoder_qkv_multiscales=[[], [], [], []], downsample_block_type="pixel_unshuffle")


def decoder_param_shapes(cfg: dict) -> Dict[str, Sequence[int]]:
    """(key -> shape) of the decoder's state dict in the reference's registration order, buffers included."""
    ch, layers = list(cfg["decoder_block_out_channels"]), list(cfg["decoder_layers_per_block"])
    n, lat, d = len(ch), cfg["latent_channels"], cfg["attention_head_dim"]
    out: Dict[str, Sequence[int]] = {}

    def bn(pre, c):
        out.update({pre + ".weight": (c,), pre + ".bias": (c,), pre + ".running_mean": (c,), pre + ".running_var": (c,),
                    pre + ".num_batches_tracked": ()})

    def rms(pre, c):
        out.update({pre + ".weight": (c,), pre + ".bias": (c,)})

    out["decoder.conv_in.weight"], out["decoder.conv_in.bias"] = (ch[-1], lat, 3, 3), (ch[-1],)
    for i in range(n):                                   # ModuleList order: up_blocks[0] .. up_blocks[n-1]
        k, c = 0, ch[i]
        if i < n - 1 and layers[i] > 0:
            pre = f"decoder.up_blocks.{i}.{k}.conv"
            out[pre + ".weight"], out[pre + ".bias"] = (4 * c, ch[i + 1], 3, 3), (4 * c,)
            k += 1
        for _ in range(layers[i]):
            pre = f"decoder.up_blocks.{i}.{k}"
            if cfg["decoder_block_types"][i] == "ResBlock":
                out[pre + ".conv1.weight"], out[pre + ".conv1.bias"], out[pre + ".conv2.weight"] = (c, c, 3, 3), (c,), (c, c, 3, 3)
                (rms if cfg["decoder_norm_types"][i] == "rms_norm" else bn)(pre + ".norm", c)
            else:
                for nm in ("to_q", "to_k", "to_v"):
                    out[pre + f".attn.{nm}.weight"] = (c, c)
                out[pre + ".attn.to_out.weight"] = (c, c)
                (rms if cfg["decoder_norm_types"][i] == "rms_norm" else bn)(pre + ".attn.norm_out", c)
                g = pre + ".conv_out"
                out[g + ".conv_inverted.weight"], out[g + ".conv_inverted.bias"] = (8 * c, c, 1, 1), (8 * c,)
                out[g + ".conv_depth.weight"], out[g + ".conv_depth.bias"] = (8 * c, 1, 3, 3), (8 * c,)
                out[g + ".conv_point.weight"] = (c, 4 * c, 1, 1)
                rms(g + ".norm", c)
            k += 1
    c0 = ch[0] if layers[0] > 0 else ch[1]
    rms("decoder.norm_out", c0)
    if layers[0] > 0:
        out["decoder.conv_out.weight"], out["decoder.conv_out.bias"] = (cfg["in_channels"], c0, 3, 3), (cfg["in_channels"],)
    else:
        out["decoder.conv_out.conv.weight"], out["decoder.conv_out.conv.bias"] = (4 * cfg["in_channels"], c0, 3, 3), \
            (4 * cfg["in_channels"],)
    return out


def seeded_weights(shapes: Dict[str, Sequence[int]], seed: int) -> Dict[str, torch.Tensor]:
    """Stand-in weights (no DC-AE checkpoint exists offline): fan-in scaled convolutions / linears, norms away from the
    identity, positive running variances."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in shapes.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(100, dtype=torch.int64)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(shape, generator=g)
        elif k.endswith("running_mean"):
            sd[k] = 0.2 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[k] = torch.randn(shape, generator=g) / fan_in ** 0.5
        elif ".norm" in k and k.endswith(".weight"):
            sd[k] = 1.0 + 0.2 * torch.randn(shape, generator=g)
        else:
            sd[k] = 0.1 * torch.randn(shape, generator=g)
    return sd
