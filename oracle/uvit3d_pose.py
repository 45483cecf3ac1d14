"""ORACLE (test infrastructure): UViT3DPose backbone forward, functional over a reference-keyed state dict.

Restates
  algorithms/dfot/backbones/u_vit/u_vit3d_pose.py:13-131   (per-pixel pose FiLM, level embeddings)
  algorithms/dfot/backbones/u_vit/u_vit3d.py:22-335         (level structure, skip connections)
  algorithms/dfot/backbones/u_vit/u_vit_blocks.py           (EmbedInput, ProjectOutput, ResBlock, NormalizeWithCond,
                                                             TransformerBlock, Downsample, Upsample)
  algorithms/dfot/backbones/modules/normalization.py:5-53   (RMSNorm), embeddings.py:390-428 (RandomDropoutPatchEmbed)
block_types must be ResBlock / TransformerBlock (the axial variant is unused by the configs), pos_emb_type "rope".
"""
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .dit3d import apply_rope, fourier_embedding, rope_angles, sinusoidal_embedding


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _rms(x, w, eps=1e-6):
    return (x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + eps)) * w


class UViT3DPoseOracle:
    def __init__(self, backbone_cfg: dict, x_shape, max_tokens: int, state_dict: Dict[str, torch.Tensor]):
        cfg = backbone_cfg
        self.sd = {k: v.detach().float() for k, v in state_dict.items()}
        self.channels = list(cfg["channels"])
        self.emb_dim = cfg["emb_channels"]
        self.p = cfg["patch_size"]
        self.block_types = list(cfg["block_types"])
        assert all(b in ("ResBlock", "TransformerBlock") for b in self.block_types)
        assert cfg["pos_emb_type"] == "rope"
        self.n_updown = list(cfg["num_updown_blocks"])
        self.n_mid = cfg["num_mid_blocks"]
        self.heads = cfg["num_heads"]
        self.T = max_tokens
        self.C, self.H, self.W = x_shape
        self.use_fourier = bool(cfg.get("use_fourier_noise_embedding", False))
        self.cond_dropout = cfg.get("external_cond_dropout", 0.0)
        self.levels = len(self.channels)
        res0 = self.W // self.p
        self.angles = {i: rope_angles(self.channels[i] // self.heads, (self.T, res0 >> i, res0 >> i))
                       for i in range(self.levels) if self.block_types[i] != "ResBlock"}

    # ---- blocks
    def res_block(self, x, emb, pre):
        sd = self.sd
        h = F.group_norm(x, 32, sd[pre + ".in_layers.0.weight"], sd[pre + ".in_layers.0.bias"], eps=1e-6)
        h = F.conv2d(F.silu(h), sd[pre + ".in_layers.2.weight"], sd[pre + ".in_layers.2.bias"], padding=1)
        e = F.conv2d(emb, sd[pre + ".emb_layer.weight"], sd[pre + ".emb_layer.bias"])
        scale, shift = e.chunk(2, dim=1)
        h = F.group_norm(h, 32, sd[pre + ".out_norm.weight"], sd[pre + ".out_norm.bias"], eps=1e-6) * (1 + scale) + shift
        h = F.conv2d(F.silu(h), sd[pre + ".out_rest.1.weight"], sd[pre + ".out_rest.1.bias"], padding=1)
        return x + h

    def transformer_block(self, x, emb, pre, level):
        sd = self.sd
        B, N, C = x.shape
        d = C // self.heads
        scale, shift = _lin(emb, sd, pre + ".norm.emb_layer").chunk(2, dim=-1)
        xn = _rms(x, sd[pre + ".norm.norm.weight"]) * (1 + scale) + shift
        fused = _lin(xn, sd, pre + ".fused_attn_mlp_proj")
        qkv, mlp_h = fused.split((3 * C, 4 * C), dim=-1)
        q, k, v = qkv.reshape(B, N, 3, self.heads, d).permute(2, 0, 3, 1, 4).unbind(0)
        q, k = _rms(q, sd[pre + ".q_norm.weight"]), _rms(k, sd[pre + ".k_norm.weight"])
        ang = self.angles[level]
        q, k = apply_rope(q, ang), apply_rope(k, ang)
        w = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1)
        a = (w @ v).transpose(1, 2).reshape(B, N, C)
        x = x + _lin(a, sd, pre + ".attn_out")
        return x + _lin(F.silu(mlp_h), sd, pre + ".mlp_out.2")

    def run_level(self, x, emb, level, prefixes):
        """x, emb: (B*T, C, h, w).  u_vit3d.py:199-282 with the pose override u_vit3d_pose.py:44-61."""
        if self.block_types[level] == "ResBlock":
            for pre in prefixes:
                x = self.res_block(x, emb, pre)
            return x
        BT, C, h, w = x.shape
        tok = lambda y: y.reshape(BT // self.T, self.T, y.shape[1], h * w).permute(0, 1, 3, 2).reshape(BT // self.T, self.T * h * w, y.shape[1])
        xt, et = tok(x), tok(emb)
        for pre in prefixes:
            xt = self.transformer_block(xt, et, pre, level)
        return xt.reshape(BT // self.T, self.T, h * w, C).permute(0, 1, 3, 2).reshape(BT, C, h, w)

    def noise_embedding(self, k):
        sd = self.sd
        e = fourier_embedding(k, sd["noise_level_pos_embedding.timesteps.freqs"],
                              sd["noise_level_pos_embedding.timesteps.phases"]) if self.use_fourier \
            else sinusoidal_embedding(k, 256)
        e = _lin(e, sd, "noise_level_pos_embedding.embedding.linear_1")
        return _lin(F.silu(e), sd, "noise_level_pos_embedding.embedding.linear_2")

    def __call__(self, x, noise_levels, external_cond, external_cond_mask: Optional[torch.Tensor] = None):
        sd = self.sd
        B, T = x.shape[:2]
        assert T == self.T and external_cond is not None
        x = F.conv2d(x.reshape(B * T, self.C, self.H, self.W).float(), sd["embed_input.proj.weight"],
                     sd["embed_input.proj.bias"], stride=self.p)
        pe = F.conv2d(external_cond.reshape(B * T, -1, self.H, self.W).float(),
                      sd["external_cond_embedding.patch_embedder.proj.weight"],
                      sd["external_cond_embedding.patch_embedder.proj.bias"], stride=self.p)
        pe = pe.reshape(B, T, *pe.shape[1:])
        if external_cond_mask is not None:        # embeddings.py:336-361: zero whole rows (eval)
            pe = torch.where(external_cond_mask.reshape(B, 1, 1, 1, 1), torch.zeros_like(pe), pe)
        emb = (self.noise_embedding(noise_levels)[..., None, None] + pe).reshape(B * T, self.emb_dim, *pe.shape[-2:])
        embs = [emb if i == 0 else F.avg_pool2d(emb, 2 ** i, 2 ** i) for i in range(self.levels)]
        before, after = [], []
        for i in range(self.levels - 1):
            x = self.run_level(x, embs[i], i, [f"down_blocks.{i}.{j}" for j in range(self.n_updown[i])])
            before.append(x)
            pre = f"down_blocks.{i}.{self.n_updown[i]}.conv"
            x = F.conv2d(F.avg_pool2d(x, 2, 2), sd[pre + ".weight"], sd[pre + ".bias"], padding=1)
            after.append(x)
        x = self.run_level(x, embs[-1], self.levels - 1, [f"mid_blocks.{j}" for j in range(self.n_mid)])
        for u in range(self.levels - 1):
            i = self.levels - 2 - u
            x = x - after.pop()
            x = F.conv2d(x, sd[f"up_blocks.{u}.0.conv.weight"], sd[f"up_blocks.{u}.0.conv.bias"], padding=1)
            x = F.interpolate(x, scale_factor=2, mode="nearest") + before.pop()
            x = self.run_level(x, embs[i], i, [f"up_blocks.{u}.{j + 1}" for j in range(self.n_updown[i])])
        x = F.conv_transpose2d(x, sd["project_output.proj.weight"], sd["project_output.proj.bias"], stride=self.p)
        return x.reshape(B, T, self.C, self.H, self.W)
