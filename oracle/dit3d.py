"""ORACLE (test infrastructure): DiT3D backbone forward — variant=full (rope_3d, learned_1d or sinusoidal_1d positions), the
factorized variants (factorized_encoder / factorized_attention: per layer a spatial block over the patches of a frame,
then a temporal block over the frames of a patch position; sinusoidal_factorized, learned_1d or sinusoidal_1d positions)
and the matrix-attention variants (full_matrix_attention / factorized_matrix_attention with matrix_block=matrix and
sinusoidal_2d positions: frames are the attention tokens, a frame's [patches, channels] matrix is projected as u^T X v) —
functional over a reference-keyed state dict.

Restates
  algorithms/dfot/backbones/dit/dit3d.py:153-192           (patchify / unpatchify)
  algorithms/dfot/backbones/dit/dit_base.py:310-425        (block loop, final layer)
  algorithms/dfot/backbones/dit/dit_blocks.py:21-44,47-123 (attention), :408-437 (AdaLN-Zero),
                                              :488-510 (block, residual-on-modulated quirk), :513-542,
                                              :211-350 (MatrixAttention), :549-652 (MatrixDiTBlock)
  algorithms/dfot/backbones/modules/embeddings.py:67-153   (noise-level embedding), :156-277 (RoPE-ND)
Third-party pieces restated from their pinned versions (see oracle/ref_shim.py):
timm PatchEmbed / Mlp, diffusers TimestepEmbedding.
"""
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


def sinusoidal_embedding(k: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """embeddings.py:112-153 with flip_sin_to_cos=True, downscale_freq_shift=0 → [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    ang = k[..., None].float() * freqs
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def fourier_embedding(k: torch.Tensor, freqs: torch.Tensor, phases: torch.Tensor) -> torch.Tensor:
    """embeddings.py:94-109."""
    y = k.to(torch.float32)[..., None] * freqs.float() + phases.float()
    return (y.cos() * math.sqrt(2)).to(k.dtype if k.is_floating_point() else torch.float32)


def rope_axis_dims(head_dim: int) -> Tuple[int, int, int]:
    """embeddings.py:254-277: rotary widths for (t, h, w)."""
    half = head_dim // 2
    q, r = divmod(half, 3)
    dims = {0: (q, q, q), 1: (q + 1, q, q), 2: (q, q + 1, q + 1)}[r]
    return tuple(2 * d for d in dims)


def rope_angles(head_dim: int, sizes: Tuple[int, int, int], theta: float = 10000.0) -> torch.Tensor:
    """embeddings.py:156-213: angle table [T*H*W, head_dim]; each frequency repeated twice."""
    T, H, W = sizes
    parts = []
    for axis, (dim, n) in enumerate(zip(rope_axis_dims(head_dim), sizes)):
        inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        ang = torch.arange(n, dtype=torch.float32)[:, None] * inv[None, :]
        ang = ang.repeat_interleave(2, dim=-1)                     # [n, dim]
        shape = [1, 1, 1, dim]
        shape[axis] = n
        parts.append(ang.reshape(shape).expand(T, H, W, dim))
    return torch.cat(parts, dim=-1).reshape(T * H * W, head_dim)


def rope_angles_1d(dim: int, n: int, theta: float = 10000.0) -> torch.Tensor:
    """embeddings.py:218-231 (RotaryEmbedding1D): angle table [n, dim]; each frequency repeated twice."""
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    return (torch.arange(n, dtype=torch.float32)[:, None] * inv[None, :]).repeat_interleave(2, dim=-1)


def apply_rope(x: torch.Tensor, angles: torch.Tensor) -> torch.Tensor:
    """embeddings.py:204-215 with interleaved rotate_half: (x0,x1) -> (-x1,x0)."""
    a = angles[: x.shape[-2]]
    pairs = x.reshape(*x.shape[:-1], -1, 2)
    rot = torch.stack((-pairs[..., 1], pairs[..., 0]), dim=-1).reshape(x.shape)
    return x * a.cos() + rot * a.sin()


def sincos_nd_table(dim: int, shape) -> torch.Tensor:
    """dit_base.py:528-580 (get_nd_sincos_pos_embed): [prod(shape), dim]; `np.meshgrid` with its default "xy" indexing, so
    for a 2-D grid the FIRST dim/2 columns encode the coordinate that varies fastest (quirk Q5)."""
    import numpy as np
    assert dim % (2 * len(shape)) == 0
    grid = np.stack(np.meshgrid(*[np.arange(n, dtype=np.float32) for n in shape]), axis=0)
    d = dim // len(shape)
    omega = 1.0 / 10000 ** (np.arange(d // 2, dtype=np.float64) / (d / 2.0))
    parts = []
    for i in range(len(shape)):
        ang = np.einsum("m,d->md", grid[i].reshape(-1), omega)
        parts.append(np.concatenate([np.sin(ang), np.cos(ang)], axis=1))
    return torch.from_numpy(np.concatenate(parts, axis=1)).float()


def _linear(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def _adaln(x, c_act, sd, prefix, chunks):
    mod = _linear(c_act, sd, prefix + ".modulation.1")
    parts = mod.chunk(chunks, dim=-1)
    y = F.layer_norm(x, (x.shape[-1],), eps=1e-6) * (1 + parts[1]) + parts[0]
    return (y, parts[2]) if chunks == 3 else y


class DiT3DOracle:
    """Callable (x[B,T,C,H,W], noise_levels[B,T], external_cond, external_cond_mask) -> like x."""

    def __init__(self, backbone_cfg: dict, x_shape, max_tokens: int, state_dict: Dict[str, torch.Tensor],
                 external_cond_dim: int = 0):
        cfg = backbone_cfg
        self.pos_emb_type = cfg.get("pos_emb_type", "rope_3d")
        self.variant = cfg.get("variant", "full")
        self.matrix = self.variant in ("full_matrix_attention", "factorized_matrix_attention")
        self.factorized = self.variant in ("factorized_encoder", "factorized_attention")
        assert self.variant == "full" or self.factorized or self.matrix
        if self.matrix:
            assert self.pos_emb_type == "sinusoidal_2d" and cfg.get("matrix_block") in ("matrix", "matrix_self", "matrix_cross")
        else:
            assert self.pos_emb_type in (("learned_1d", "sinusoidal_1d", "sinusoidal_factorized") if self.factorized
                                         else ("rope_3d", "learned_1d", "sinusoidal_1d"))
        self.sd = {k: v.detach().float() for k, v in state_dict.items()}
        self.p = cfg["patch_size"]
        self.C, self.H, self.W = x_shape
        self.gh, self.gw = self.H // self.p, self.W // self.p
        self.P = self.gh * self.gw
        self.D = cfg["embed_row_dim"] if self.matrix else cfg["hidden_size"]      # dit3d.py:113-118
        self.depth = cfg["depth"]
        self.heads = cfg.get("num_heads") or 1
        self.dh = self.D // self.heads
        if self.matrix:                                                            # dit_base.py:129-149, 296-308
            self.col_heads, self.row_heads = cfg["num_col_heads"], cfg["num_row_heads"]
            self.hc, self.hr = cfg["embed_col_dim"] // self.col_heads, cfg["embed_row_dim"] // self.row_heads
            self.flatten_rope, self.multi_token = bool(cfg.get("flatten_matrix_rope")), bool(cfg.get("matrix_multi_token"))
            self.fixed_u = cfg.get("fixed_u")
            self.matrix_angles = None
            if cfg.get("use_temporal_rope"):
                self.matrix_angles = rope_angles_1d(self.hc * self.hr if self.flatten_rope else self.hr, max_tokens)
        self.use_fourier = bool(cfg.get("use_fourier_noise_embedding", False))
        self.external_cond_dim = external_cond_dim
        self.cond_dropout = cfg.get("external_cond_dropout", 0.0)
        # dit_base.py:230-253: rope_3d rotates q / k in every block; learned_1d / sinusoidal_1d add one table to the tokens
        self.angles = rope_angles(self.dh, (max_tokens, self.gh, self.gw)) if self.pos_emb_type == "rope_3d" else None
        if self.pos_emb_type == "learned_1d":
            self.pos_emb = self.sd["dit_base.pos_emb.pos_emb"]
        elif self.pos_emb_type == "sinusoidal_1d":      # non-persistent buffer (dit_base.py:514-521, 528-580)
            import numpy as np
            n = max_tokens * self.P
            omega = 1.0 / 10000 ** (np.arange(self.D // 2, dtype=np.float64) / (self.D / 2.0))
            ang = np.einsum("m,d->md", np.arange(n, dtype=np.float32).astype(np.float64), omega)
            self.pos_emb = torch.from_numpy(np.concatenate([np.sin(ang), np.cos(ang)], axis=1)).float().unsqueeze(0)
        if self.pos_emb_type == "sinusoidal_factorized":   # dit_base.py:265-274: 2-D spatial table + 1-D temporal table
            self.spatial_pos = sincos_nd_table(self.D, (self.gh, self.gw)).unsqueeze(0)
            self.temporal_pos = sincos_nd_table(self.D, (max_tokens,)).unsqueeze(0)
        if self.pos_emb_type == "sinusoidal_2d":           # dit_base.py:254-258: the spatial table, added per frame (:356-362)
            self.spatial_pos = sincos_nd_table(self.D, (self.gh, self.gw)).unsqueeze(0)
        # dit_base.py:185,192: MLP exists only if spatial_mlp_ratio is set (fork quirk Q2)
        self.use_mlp = "dit_base.blocks.0.mlp.fc1.weight" in self.sd
        self.taps = None  # optional dict filled with intermediates for kernel-level parity tests

    def noise_embedding(self, k: torch.Tensor) -> torch.Tensor:
        sd = self.sd
        if self.use_fourier:
            e = fourier_embedding(k, sd["noise_level_pos_embedding.timesteps.freqs"],
                                  sd["noise_level_pos_embedding.timesteps.phases"])
        else:
            e = sinusoidal_embedding(k, 256)
        e = _linear(e, sd, "noise_level_pos_embedding.embedding.linear_1")
        return _linear(F.silu(e), sd, "noise_level_pos_embedding.embedding.linear_2")

    def cond_embedding(self, cond: torch.Tensor, cond_mask: Optional[torch.Tensor]) -> torch.Tensor:
        # embeddings.py:364-387: dropout_prob == 0 → plain TimestepEmbedding that ignores the mask
        sd = self.sd
        if self.cond_dropout == 0:
            e = _linear(cond, sd, "external_cond_embedding.linear_1")
            return _linear(F.silu(e), sd, "external_cond_embedding.linear_2")
        e = _linear(cond, sd, "external_cond_embedding.embedding.linear_1")
        e = _linear(F.silu(e), sd, "external_cond_embedding.embedding.linear_2")
        if cond_mask is not None:
            e = torch.where(cond_mask.reshape(-1, *([1] * (e.ndim - 1))), torch.zeros_like(e), e)
        return e

    def attention(self, y: torch.Tensor, i: int, group: str = "blocks") -> torch.Tensor:
        B, N, D = y.shape
        qkv = _linear(y, self.sd, f"dit_base.{group}.{i}.attn.qkv")
        q, k, v = qkv.reshape(B, N, 3, self.heads, self.dh).permute(2, 0, 3, 1, 4).unbind(0)
        if self.angles is not None:
            q, k = apply_rope(q, self.angles), apply_rope(k, self.angles)
        w = torch.softmax(q @ k.transpose(-2, -1) * (1 / math.sqrt(self.dh)), dim=-1)
        o = (w @ v).transpose(1, 2).reshape(B, N, D)
        if self.taps is not None and i == 0:
            self.taps.update(q0=q, k0=k, v0=v, attn0=o)
        return _linear(o, self.sd, f"dit_base.{group}.{i}.attn.proj")

    def matrix_attention(self, y: torch.Tensor, pre: str) -> torch.Tensor:
        """dit_blocks.py:289-350.  y [B, L, N, D]: L frames (the attention tokens) of N patch rows; heads = col x row heads,
        a head's feature is its [hc, hr] sub-matrix."""
        sd = self.sd
        B, L, N, D = y.shape
        eye = torch.eye(N) if self.fixed_u == "identity" else None          # :267-269 (plain tensors, not parameters)
        qkv = torch.einsum("nm,blnd,dk->blmk", eye if eye is not None else sd[pre + ".qkv_u"], y, sd[pre + ".qkv_v"])
        if pre + ".qkv_bias" in sd:
            qkv = qkv + sd[pre + ".qkv_bias"]
        C, R, hc, hr = self.col_heads, self.row_heads, self.hc, self.hr
        q, k, v = qkv.reshape(B, L, C, hc, 3, R, hr).permute(4, 0, 2, 5, 1, 3, 6).unbind(0)   # b c r l n d
        if self.matrix_angles is not None:
            if self.flatten_rope:                                           # rotate the flattened (n d) feature over l
                q = apply_rope(q.reshape(B, C, R, L, hc * hr), self.matrix_angles).reshape(q.shape)
                k = apply_rope(k.reshape(B, C, R, L, hc * hr), self.matrix_angles).reshape(k.shape)
            else:                                                           # every row n rotated over l with the same table
                q = apply_rope(q.transpose(3, 4), self.matrix_angles).transpose(3, 4)
                k = apply_rope(k.transpose(3, 4), self.matrix_angles).transpose(3, 4)
        if self.multi_token:                                                # :324-332: one softmax per row n
            w = torch.softmax(torch.einsum("bcrlnd,bcrknd->bcrnlk", q * hr ** -0.5, k), dim=-1)
            o = torch.einsum("bcrnlk,bcrknd->bcrlnd", w, v)
        else:                                                               # :333-338: the whole sub-matrix is the feature
            w = torch.softmax(torch.einsum("bcrlnd,bcrknd->bcrlk", q * (hc * hr) ** -0.5, k), dim=-1)
            o = torch.einsum("bcrlk,bcrknd->bcrlnd", w, v)
        o = o.permute(0, 3, 1, 4, 2, 5).reshape(B, L, C * hc, R * hr)       # b l (c n) (r d)
        o = torch.einsum("nm,blnd,dk->blmk", eye if eye is not None else sd[pre + ".proj_u"], o, sd[pre + ".proj_v"])
        if pre + ".proj_bias" in sd:
            o = o + sd[pre + ".proj_bias"]
        return o

    def token_attention(self, y: torch.Tensor, pre: str, heads: int, kv_src: Optional[torch.Tensor] = None) -> torch.Tensor:
        """dit_blocks.py:81-123 (Attention) / :162-208 (CrossAttention: q from y, k / v from kv_src) without RoPE, before the
        output projection's residual: returns proj(softmax(q k^T / sqrt(d)) v)."""
        sd = self.sd
        B, N, D = y.shape
        dh = D // heads
        if kv_src is None:
            q, k, v = _linear(y, sd, pre + ".qkv").reshape(B, N, 3, heads, dh).permute(2, 0, 3, 1, 4).unbind(0)
        else:
            q = _linear(y, sd, pre + ".q_proj").reshape(B, N, heads, dh).permute(0, 2, 1, 3)
            k, v = _linear(kv_src, sd, pre + ".kv_proj").reshape(B, N, 2, heads, dh).permute(2, 0, 3, 1, 4).unbind(0)
        w = torch.softmax(q @ k.transpose(-2, -1) * (1 / math.sqrt(dh)), dim=-1)
        return _linear((w @ v).transpose(1, 2).reshape(B, N, D), sd, pre + ".proj")

    def block(self, h, c_act, i: int, group: str, n_frames: int = 0):
        """dit_blocks.py:488-510 (the MLP exists iff the block was built with a positive ratio); :626-652 for a
        MatrixDiTBlock (same block around MatrixAttention over the n_frames frames of the row); :734-769 MatrixCrossDiTBlock
        (per frame, the modulated tokens attend to the matrix attention's output); :852-883 MatrixSelfDiTBlock (matrix
        attention, then self-attention inside every frame, then the MLP)."""
        sd, pre = self.sd, f"dit_base.{group}.{i}"
        y, gate = _adaln(h, c_act, sd, pre + ".norm1", 3)
        B, N, D = y.shape
        mlp_norm = ".norm2"
        if pre + ".attn1.qkv_v" in sd:                               # matrix_self / matrix_cross (attn1 never has a bias)
            T, P = n_frames, N // n_frames
            x1 = self.matrix_attention(y.reshape(B, T, P, D), pre + ".attn1").reshape(B, N, D)
            mlp_norm = ".norm3"
            if pre + ".attn2.q_proj.weight" in sd:                   # cross: q from y, k / v from x1, inside every frame
                att = self.token_attention(y.reshape(B * T, P, D), pre + ".attn2", self.row_heads,
                                           x1.reshape(B * T, P, D)).reshape(B, N, D)
                h = y + gate * att
            else:                                                    # self: two gated halves
                h = y + gate * x1
                y2, gate2 = _adaln(h, c_act, sd, pre + ".norm2", 3)
                att = self.token_attention(y2.reshape(B * T, P, D), pre + ".attn2", self.row_heads).reshape(B, N, D)
                h = y2 + gate2 * att
        elif pre + ".attn.qkv_v" in sd:
            att = self.matrix_attention(y.reshape(B, n_frames, N // n_frames, D), pre + ".attn").reshape(B, N, D)
            h = y + gate * att
        else:
            h = y + gate * self.attention(y, i, group)               # residual base is the modulated tensor (Q1)
        if pre + ".mlp.fc1.weight" in sd:
            z, gate2 = _adaln(h, c_act, sd, pre + mlp_norm, 3)
            m = _linear(F.gelu(_linear(z, sd, pre + ".mlp.fc1"), approximate="tanh"), sd, pre + ".mlp.fc2")
            h = z + gate2 * m
        return h

    def factorized_blocks(self, tok, c_act, B: int, T: int):
        """dit_base.py:355-412 for the factorized variants: tokens (b, t, p); every layer runs a spatial block on
        ((b t), p) and a temporal block on ((b p), t); the temporal table is added before the first temporal block."""
        P, D = self.P, self.D
        h, c = tok.reshape(B, T, P, D), c_act.reshape(B, T, P, D)
        if self.pos_emb_type == "sinusoidal_factorized":
            h = h + self.spatial_pos[:, :P].reshape(1, 1, P, D)
        for i in range(self.depth):
            h = self.block(h.reshape(B * T, P, D), c.reshape(B * T, P, D), i, "blocks").reshape(B, T, P, D)
            ht, ct = h.transpose(1, 2).reshape(B * P, T, D), c.transpose(1, 2).reshape(B * P, T, D)
            if i == 0 and self.pos_emb_type == "sinusoidal_factorized":
                ht = ht + self.temporal_pos[:, :T]
            ht = self.block(ht, ct, i, "temporal_blocks")
            h = ht.reshape(B, P, T, D).transpose(1, 2)
        return h.reshape(B, T * P, D)

    def matrix_blocks(self, tok, c_act, B: int, T: int):
        """dit_base.py:355-416 for the matrix variants: sinusoidal_2d table per frame; full_matrix_attention runs one
        MatrixDiTBlock per layer on the (t p) tokens, factorized_matrix_attention a spatial DiTBlock on ((b t), p) and
        then a MatrixDiTBlock (`temporal_blocks`) on (b, (t p))."""
        P, D = self.P, self.D
        h = (tok.reshape(B, T, P, D) + self.spatial_pos[:, :P].reshape(1, 1, P, D)).reshape(B, T * P, D)
        for i in range(self.depth):
            if self.variant == "full_matrix_attention":
                h = self.block(h, c_act, i, "blocks", T)
            else:
                h = self.block(h.reshape(B * T, P, D), c_act.reshape(B * T, P, D), i, "blocks").reshape(B, T * P, D)
                h = self.block(h, c_act, i, "temporal_blocks", T)
        return h

    def __call__(self, x, noise_levels, external_cond=None, external_cond_mask=None):
        B, T = x.shape[:2]
        sd = self.sd
        tok = F.conv2d(x.reshape(B * T, self.C, self.H, self.W).float(), sd["patch_embedder.proj.weight"],
                       sd["patch_embedder.proj.bias"], stride=self.p)
        tok = tok.flatten(2).transpose(1, 2).reshape(B, T * self.P, self.D)
        if self.angles is None and self.pos_emb_type not in ("sinusoidal_factorized", "sinusoidal_2d"):
            tok = tok + self.pos_emb[:, : tok.shape[1]]             # dit_base.py:352-353, 523-525
        emb = self.noise_embedding(noise_levels)
        if external_cond is not None:
            if "external_cond_embedding.embedding_table.weight" in sd:   # label conditioning (dit3d.py:171-173): a table
                emb = emb + sd["external_cond_embedding.embedding_table.weight"][external_cond.long().reshape(B, -1)]
            else:
                emb = emb + self.cond_embedding(external_cond.float(), external_cond_mask)
        c_act = F.silu(emb.repeat_interleave(self.P, dim=1))      # per-token copy of a per-frame vector
        h = tok
        for i in range(0 if self.factorized or self.matrix else self.depth):
            pre = f"dit_base.blocks.{i}"
            y, gate = _adaln(h, c_act, sd, pre + ".norm1", 3)
            h = y + gate * self.attention(y, i)                    # residual base is the modulated tensor (Q1)
            if self.use_mlp:
                z, gate2 = _adaln(h, c_act, sd, pre + ".norm2", 3)
                m = _linear(F.gelu(_linear(z, sd, pre + ".mlp.fc1"), approximate="tanh"), sd, pre + ".mlp.fc2")
                h = z + gate2 * m
            if self.taps is not None and i == 0:
                self.taps.update(block0=h)
        if self.factorized:
            h = self.factorized_blocks(tok, c_act, B, T)
        if self.matrix:
            h = self.matrix_blocks(tok, c_act, B, T)
        h = _adaln(h, c_act, sd, "dit_base.final_layer.norm_final", 2)
        h = _linear(h, sd, "dit_base.final_layer.linear")           # [B, T*P, p*p*C]
        h = h.reshape(B, T, self.gh, self.gw, self.p, self.p, self.C)
        return h.permute(0, 1, 6, 2, 4, 3, 5).reshape(B, T, self.C, self.H, self.W)
