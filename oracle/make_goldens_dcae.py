"""TEST INFRASTRUCTURE — fixture of the DC-AE decoder by EXECUTING the reference's model code (authoring container only):
    python -m oracle.make_goldens_dcae
`algorithms/vae/dc_ae/autoencoder_dc_model.py` is imported as it is; the modules it pulls from diffusers==0.32.2 (not
installed here, requirements.txt:4) are replaced by stand-ins that restate their published behaviour: RMSNorm,
get_normalization, get_activation, GLUMBConv, SanaMultiscaleAttnProcessor2_0 — so the fixture pins the reference's OWN
code (Decoder, DCUpBlock2d, ResBlock, EfficientViTBlock, apply_linear_attention) exactly and the diffusers pieces only as
far as the restatement is right ("parity unpinned" for those, see oracle/dc_ae.py).
Writes tests/golden/dcae_decode.{npz,json}: weights are regenerated from the seed by oracle.dc_ae.seeded_weights."""
import importlib.util
import json
import numbers
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.dc_ae import DCAEDecoderOracle, decoder_param_shapes, seeded_weights, small_cfg  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED, DATA_SEED = 21, 22


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def install_diffusers_standins():
    for n in ["diffusers", "diffusers.models", "diffusers.utils", "diffusers.models.autoencoders", "diffusers.models.transformers"]:
        if n not in sys.modules:
            _mod(n)
    cu = _mod("diffusers.configuration_utils")
    cu.ConfigMixin = type("ConfigMixin", (), {})
    cu.register_to_config = lambda f: f
    act = _mod("diffusers.models.activations")
    act.get_activation = lambda name: {"relu": nn.ReLU, "silu": nn.SiLU, "swish": nn.SiLU, "relu6": nn.ReLU6, "gelu": nn.GELU}[name]()
    vae = _mod("diffusers.models.autoencoders.vae")
    vae.DecoderOutput = vae.EncoderOutput = type("Output", (), {"__init__": lambda self, **kw: self.__dict__.update(kw)})
    mu = _mod("diffusers.models.modeling_utils")
    mu.ModelMixin = nn.Module
    au = _mod("diffusers.utils.accelerate_utils")
    au.apply_forward_hook = lambda f: f

    class RMSNorm(nn.Module):                            # diffusers/models/normalization.py (0.32.2)
        def __init__(self, dim, eps: float, elementwise_affine: bool = True, bias: bool = False):
            super().__init__()
            self.eps = eps
            dim = (dim,) if isinstance(dim, numbers.Integral) else dim
            self.weight = nn.Parameter(torch.ones(dim)) if elementwise_affine else None
            self.bias = nn.Parameter(torch.zeros(dim)) if elementwise_affine and bias else None

        def forward(self, x):
            var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
            x = x * torch.rsqrt(var + self.eps)
            if self.weight is not None:
                x = x * self.weight
                if self.bias is not None:
                    x = x + self.bias
            return x

    def get_normalization(norm_type="batch_norm", num_features=None, eps=1e-5, elementwise_affine=True, bias=True):
        if norm_type == "rms_norm":
            return RMSNorm(num_features, eps=eps, elementwise_affine=elementwise_affine, bias=bias)
        if norm_type == "batch_norm":
            return nn.BatchNorm2d(num_features, eps=eps, affine=elementwise_affine)
        raise ValueError(norm_type)

    nm = _mod("diffusers.models.normalization")
    nm.RMSNorm, nm.get_normalization = RMSNorm, get_normalization

    class GLUMBConv(nn.Module):                          # diffusers/models/transformers/sana_transformer.py (0.32.2)
        def __init__(self, in_channels, out_channels, expand_ratio=4, norm_type=None, residual_connection=True):
            super().__init__()
            hidden = int(expand_ratio * in_channels)
            self.norm_type, self.residual_connection = norm_type, residual_connection
            self.nonlinearity = nn.SiLU()
            self.conv_inverted = nn.Conv2d(in_channels, hidden * 2, 1, 1, 0)
            self.conv_depth = nn.Conv2d(hidden * 2, hidden * 2, 3, 1, 1, groups=hidden * 2)
            self.conv_point = nn.Conv2d(hidden, out_channels, 1, 1, 0, bias=False)
            self.norm = RMSNorm(out_channels, eps=1e-5, elementwise_affine=True, bias=True) if norm_type == "rms_norm" else None

        def forward(self, h):
            residual = h
            h = self.conv_depth(self.nonlinearity(self.conv_inverted(h)))
            h, gate = torch.chunk(h, 2, dim=1)
            h = self.conv_point(h * self.nonlinearity(gate))
            if self.norm_type == "rms_norm":
                h = self.norm(h.movedim(1, -1)).movedim(-1, 1)
            return h + residual if self.residual_connection else h

    st = _mod("diffusers.models.transformers.sana_transformer")
    st.GLUMBConv = GLUMBConv

    class SanaMultiscaleAttentionProjection(nn.Module):  # only built for non-empty qkv_multiscales (unused by the tree)
        def __init__(self, in_channels, num_attention_heads, kernel_size):
            super().__init__()
            ch = 3 * in_channels
            self.proj_in = nn.Conv2d(ch, ch, kernel_size, padding=kernel_size // 2, groups=ch, bias=False)
            self.proj_out = nn.Conv2d(ch, ch, 1, 1, 0, groups=3 * num_attention_heads, bias=False)

        def forward(self, h):
            return self.proj_out(self.proj_in(h))

    class SanaMultiscaleAttnProcessor2_0:                # diffusers/models/attention_processor.py (0.32.2)
        def __call__(self, attn, hidden_states):
            height, width = hidden_states.shape[-2:]
            use_linear = height * width > attn.attention_head_dim
            residual = hidden_states
            batch_size = hidden_states.shape[0]
            original_dtype = hidden_states.dtype
            hidden_states = hidden_states.movedim(1, -1)
            hidden_states = torch.cat([attn.to_q(hidden_states), attn.to_k(hidden_states), attn.to_v(hidden_states)], dim=3)
            hidden_states = hidden_states.movedim(-1, 1)
            multi = [hidden_states] + [block(hidden_states) for block in attn.to_qkv_multiscale]
            hidden_states = torch.cat(multi, dim=1)
            if use_linear:
                hidden_states = hidden_states.to(dtype=torch.float32)
            hidden_states = hidden_states.reshape(batch_size, -1, 3 * attn.attention_head_dim, height * width)
            query, key, value = hidden_states.chunk(3, dim=2)
            query, key = attn.nonlinearity(query), attn.nonlinearity(key)
            if use_linear:
                hidden_states = attn.apply_linear_attention(query, key, value).to(dtype=original_dtype)
            else:
                hidden_states = attn.apply_quadratic_attention(query, key, value)
            hidden_states = torch.reshape(hidden_states, (batch_size, -1, height, width))
            hidden_states = attn.to_out(hidden_states.movedim(1, -1)).movedim(-1, 1)
            if attn.norm_type == "rms_norm":
                hidden_states = attn.norm_out(hidden_states.movedim(1, -1)).movedim(-1, 1)
            else:
                hidden_states = attn.norm_out(hidden_states)
            return hidden_states + residual if attn.residual_connection else hidden_states

    ap = _mod("diffusers.models.attention_processor")
    ap.SanaMultiscaleAttentionProjection, ap.SanaMultiscaleAttnProcessor2_0 = SanaMultiscaleAttentionProjection, \
        SanaMultiscaleAttnProcessor2_0
    # reference-side utility modules the file imports but the decoder never calls
    su = _mod("utils.storage_utils")
    su.safe_torch_save = lambda *a, **k: None
    ck = _mod("utils.ckpt_utils")
    for n in ["is_wandb_run_path", "is_hf_path"]:
        setattr(ck, n, lambda p: False)
    ck.wandb_to_local_path = ck.download_pretrained = lambda p: p


def reference_module():
    ref_shim.install()
    install_diffusers_standins()
    path = os.path.join(ref_shim.REF, "algorithms", "vae", "dc_ae", "autoencoder_dc_model.py")
    spec = importlib.util.spec_from_file_location("ref_autoencoder_dc_model", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    mod = reference_module()
    cfg = small_cfg()
    model = mod.MyAutoencoderDC(ref_shim.to_dc(cfg)).eval()
    shapes = decoder_param_shapes(cfg)
    ref_keys = [k for k in model.state_dict().keys() if k.startswith("decoder.")]
    assert ref_keys == list(shapes.keys()), [(a, b) for a, b in zip(ref_keys, shapes) if a != b][:5]
    for k, v in model.state_dict().items():
        if k.startswith("decoder."):
            assert tuple(v.shape) == tuple(shapes[k]), (k, v.shape, shapes[k])
    sd = seeded_weights(shapes, WEIGHT_SEED)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("encoder.") for k in missing)
    g = torch.Generator().manual_seed(DATA_SEED)
    z = torch.randn((3, cfg["latent_channels"], 8, 8), generator=g)   # 64 tokens > head_dim: the linear-attention branch, as at DMLab size
    with torch.no_grad():
        out = model.decode(z)
        mine = DCAEDecoderOracle(sd, cfg).decode(z)
    print("reference decode", tuple(out.shape), "max |oracle - reference| =", (mine - out).abs().max().item())
    np.savez_compressed(os.path.join(OUT, "dcae_decode.npz"), z=z.numpy(), image=out.numpy())
    with open(os.path.join(OUT, "dcae_decode.json"), "w") as f:
        json.dump(dict(cfg=cfg, weight_seed=WEIGHT_SEED, data_seed=DATA_SEED, keys=ref_keys), f, indent=1)


if __name__ == "__main__":
    main()
