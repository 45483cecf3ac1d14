"""DiT3D backbone on hand-written sm_100a kernels: variant=full (pos_emb_type rope_3d — the shipped dit3d.yaml — learned_1d,
sinusoidal_1d), the factorized variants (factorized_encoder / factorized_attention, dit3d_factorized_attention.yaml:
per layer a spatial block over the patches of a frame and a temporal block over the frames of a patch position) and the
matrix-attention variants (full_matrix_attention / factorized_matrix_attention, dit3d_full_matrix.yaml /
dit3d_factorized_matrix.yaml: frames are the attention tokens of a MatrixDiTBlock; matrix_block = matrix | matrix_self |
matrix_cross, the latter two with token attention inside every frame behind the matrix attention).

Drop-in for the reference class
    algorithms/dfot/backbones/dit/dit3d.py:11-192  (DiT3D),
    algorithms/dfot/backbones/dit/dit_base.py:77-425 (DiTBase), dit_blocks.py:378-542 (blocks),
    dit_blocks.py:211-350 (MatrixAttention), :549-652 (MatrixDiTBlock), :655-883 (MatrixCrossDiTBlock, MatrixSelfDiTBlock)
same constructor signature, same ``forward(x, noise_levels, external_cond, external_cond_mask)`` and the
same ``state_dict()`` keys, so a reference checkpoint loads unchanged.  Parameters are kept in fp32
under the reference's names; kernel-layout copies (bf16, concatenated modulation weights, padded
patch-embed / condition weights, RoPE cos/sin table) are (re)packed lazily whenever a parameter changes.

Per forward (R rows, T frames, P patches/frame, M = R*T*P tokens, D hidden):
    K4/patchify → [tcgen05 GEMM] patch-embed → x f32
    noise features → GEMM(+SiLU) → GEMM → (+cond emb) SiLU → ONE GEMM for the adaLN modulation of all blocks
        (per *frame*, M = R*T rows — the reference recomputes it per token: 30 % of its FLOPs at K600-XL)
    per block: K1 adaLN-LN → GEMM qkv (+bias, RoPE-3D, q-scale fused in the epilogue) → K3 attention →
               GEMM proj (+bias, gate, residual fused) → K1 → GEMM fc1 (+bias, GELU) → GEMM fc2 (+gate, residual)
    K1 (final) → GEMM final → unpatchify
Factorized variants reuse the same kernels: a spatial block is the block above with R*T "samples" of P tokens; for a
temporal block the fp32 token stream is transposed to (row, patch, frame) order (data movement), the block runs with
R*P "samples" of T tokens and a per-token copy of its modulation columns (the frame of a token is no longer m / P), and
the stream is transposed back.
A matrix block keeps the (row, frame, patch) token order and per-frame modulation of a full block; its attention is
    K1 adaLN-LN → patch-mix (the `qkv_u` factor: P patch rows of a frame → Mc rows) → GEMM qkv_v on R*Mc*T rows (+ 1-D RoPE
    over the frame index, q-scale) → K3 attention over the T frames (R*Mc sequences) → GEMM proj_v → patch-expand (the
    `proj_u` factor back to P rows, + proj_bias, gate, residual)
i.e. the u factor is contracted first (the reference's einsum 'nm,blnd,dk->blmk' leaves the order open), so the GEMMs see
Mc rows per frame instead of P.
"""
import math
import os
from typing import Optional

import numpy as np
import torch
from torch import nn

from dfot_b200 import _abi, ops
from dfot_b200.config import to_config

LOG2E = 1.4426950408889634


# ------------------------------------------------------------------ parameter containers (reference key names)
class _TimestepMLP(nn.Module):          # diffusers TimestepEmbedding: linear_1 → SiLU → linear_2
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)


class _LabelEmbedding(nn.Module):       # diffusers LabelEmbedding (base_backbone.py:47-51): eval = a table lookup; the
    def __init__(self, num_classes: int, dim: int, dropout_prob: float):    # extra row is the training-time "dropped" class
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + int(dropout_prob > 0), dim)
        self.num_classes, self.dropout_prob = num_classes, dropout_prob


class _Fourier(nn.Module):              # embeddings.py:94-109 (persistent random buffers)
    def __init__(self, dim: int):
        super().__init__()
        self.register_buffer("freqs", 2 * math.pi * torch.randn(dim))
        self.register_buffer("phases", 2 * math.pi * torch.rand(dim))


class _NoiseLevelEmbedding(nn.Module):  # embeddings.py:67-91
    def __init__(self, dim: int, emb_dim: int, use_fourier: bool):
        super().__init__()
        if use_fourier:
            self.timesteps = _Fourier(dim)
        self.embedding = _TimestepMLP(dim, emb_dim)


class _CondEmbeddingDropout(nn.Module):  # embeddings.py:364-387 with dropout_prob > 0
    def __init__(self, cond_dim: int, emb_dim: int):
        super().__init__()
        self.embedding = _TimestepMLP(cond_dim, emb_dim)


class _PatchEmbed(nn.Module):           # timm PatchEmbed: Conv2d(k = s = p)
    def __init__(self, in_chans: int, dim: int, p: int):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=p, stride=p, bias=True)


class _AdaLN(nn.Module):                # dit_blocks.py:378-437
    def __init__(self, dim: int, chunks: int):
        super().__init__()
        self.modulation = nn.Sequential(nn.SiLU(), nn.Linear(dim, chunks * dim, bias=True))
        nn.init.zeros_(self.modulation[-1].weight)
        nn.init.zeros_(self.modulation[-1].bias)


class _Attention(nn.Module):            # dit_blocks.py:47-79
    def __init__(self, dim: int):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):                  # timm Mlp
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):                # dit_blocks.py:440-510
    def __init__(self, dim: int, mlp_ratio: Optional[float]):
        super().__init__()
        self.norm1 = _AdaLN(dim, 3)
        self.attn = _Attention(dim)
        self.use_mlp = mlp_ratio is not None and mlp_ratio > 0.0
        if self.use_mlp:
            self.norm2 = _AdaLN(dim, 3)
            self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        for lin in [self.attn.qkv, self.attn.proj] + ([self.mlp.fc1, self.mlp.fc2] if self.use_mlp else []):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)

    def adaln_norms(self):              # the block's AdaLN-Zero layers in execution order
        return [self.norm1] + ([self.norm2] if self.use_mlp else [])


class _MatrixAttention(nn.Module):      # dit_blocks.py:215-287 (parameters in the reference's registration order)
    def __init__(self, col_dim: int, row_dim: int, embed_col_dim: int, embed_row_dim: int, use_bias: bool,
                 fixed_u: Optional[str]):
        super().__init__()
        self.fixed_u = fixed_u
        if fixed_u is None:
            self.qkv_u = nn.Parameter(torch.empty(col_dim, embed_col_dim))
            self.proj_u = nn.Parameter(torch.empty(embed_col_dim, col_dim))
        elif fixed_u == "identity":      # plain tensors in the reference (:267-269), not part of the state dict
            if embed_col_dim != col_dim:
                raise ValueError("fixed_u='identity' needs embed_col_dim == the number of patches per frame")
        else:
            raise ValueError(f"Invalid fixed_u value: {fixed_u}. It should be 'identity', None")
        self.qkv_v = nn.Parameter(torch.empty(row_dim, 3 * embed_row_dim))
        self.proj_v = nn.Parameter(torch.empty(embed_row_dim, row_dim))
        if use_bias:
            self.qkv_bias = nn.Parameter(torch.zeros(embed_col_dim, 3 * embed_row_dim))
            self.proj_bias = nn.Parameter(torch.zeros(col_dim, row_dim))
        for w in [self.qkv_v, self.proj_v] + ([self.qkv_u, self.proj_u] if fixed_u is None else []):
            nn.init.xavier_uniform_(w)                   # dit_blocks.py:604-621


class _MatrixBlock(nn.Module):          # dit_blocks.py:549-652
    def __init__(self, col_dim: int, dim: int, embed_col_dim: int, mlp_ratio: Optional[float], use_bias: bool,
                 fixed_u: Optional[str]):
        super().__init__()
        self.norm1 = _AdaLN(dim, 3)
        self.attn = _MatrixAttention(col_dim, dim, embed_col_dim, dim, use_bias, fixed_u)
        self.use_mlp = mlp_ratio is not None and mlp_ratio > 0.0
        if self.use_mlp:
            self.norm2 = _AdaLN(dim, 3)
            self.mlp = _Mlp(dim, int(dim * mlp_ratio))
            for lin in (self.mlp.fc1, self.mlp.fc2):
                nn.init.xavier_uniform_(lin.weight)
                nn.init.zeros_(lin.bias)

    def adaln_norms(self):
        return [self.norm1] + ([self.norm2] if self.use_mlp else [])


class _CrossAttention(nn.Module):       # dit_blocks.py:125-160 (q from the tokens, k / v from a second stream)
    def __init__(self, dim: int):
        super().__init__()
        self.q_proj = nn.Linear(dim, dim, bias=True)
        self.kv_proj = nn.Linear(dim, 2 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)


class _MatrixTokenBlock(nn.Module):     # dit_blocks.py:655-769 (MatrixCrossDiTBlock), :772-883 (MatrixSelfDiTBlock)
    """Matrix attention over the frames (attn1: never a bias or a fixed u — the two blocks drop those keyword arguments)
    followed by token attention inside every frame (attn2: self-attention behind a second AdaLN, or cross-attention of the
    modulated tokens to attn1's output).  Modules in the reference's registration order."""

    def __init__(self, col_dim: int, dim: int, embed_col_dim: int, mlp_ratio: Optional[float], cross: bool):
        super().__init__()
        self.cross = cross
        self.norm1 = _AdaLN(dim, 3)
        self.attn1 = _MatrixAttention(col_dim, dim, embed_col_dim, dim, False, None)
        if cross:
            self.attn2 = _CrossAttention(dim)
            lins = [self.attn2.q_proj, self.attn2.kv_proj, self.attn2.proj]
        else:
            self.norm2 = _AdaLN(dim, 3)
            self.attn2 = _Attention(dim)
            lins = [self.attn2.qkv, self.attn2.proj]
        self.use_mlp = mlp_ratio is not None                  # dit_blocks.py:702, :819 (not "> 0" as in DiTBlock)
        if self.use_mlp:
            self.norm3 = _AdaLN(dim, 3)
            self.mlp = _Mlp(dim, int(dim * mlp_ratio))
            lins += [self.mlp.fc1, self.mlp.fc2]
        for lin in lins:
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)

    def adaln_norms(self):
        return [self.norm1] + ([] if self.cross else [self.norm2]) + ([self.norm3] if self.use_mlp else [])


class _FinalLayer(nn.Module):           # dit_blocks.py:513-542
    def __init__(self, dim: int, out_channels: int):
        super().__init__()
        self.norm_final = _AdaLN(dim, 2)
        self.linear = nn.Linear(dim, out_channels, bias=True)
        nn.init.zeros_(self.linear.weight)
        nn.init.zeros_(self.linear.bias)


def sincos_1d_table(dim: int, n: int) -> torch.Tensor:
    """dit_base.py:528-580 for a 1-D shape: [n, dim] = [sin(pos * w) | cos(pos * w)], w_i = 10000^(-2i/dim), float64 → f32."""
    omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
    out = np.einsum("m,d->md", np.arange(n, dtype=np.float32).astype(np.float64), omega)
    return torch.from_numpy(np.concatenate([np.sin(out), np.cos(out)], axis=1)).float()


def sincos_nd_table(dim: int, shape) -> torch.Tensor:
    """dit_base.py:528-580 for an n-D grid: the per-axis tables side by side, dim/n columns each.  The reference builds the
    grid with `np.meshgrid` and its default "xy" indexing — the first block of columns encodes the coordinate that varies
    fastest along the flattened grid (quirk Q5) — so the same call is made here."""
    assert dim % (2 * len(shape)) == 0
    grid = np.meshgrid(*[np.arange(n, dtype=np.float32) for n in shape])
    return torch.cat([sincos_1d_table_at(dim // len(shape), g.reshape(-1)) for g in grid], dim=1)


def sincos_1d_table_at(dim: int, pos: np.ndarray) -> torch.Tensor:
    omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
    out = np.einsum("m,d->md", pos.astype(np.float64), omega)
    return torch.from_numpy(np.concatenate([np.sin(out), np.cos(out)], axis=1)).float()


class _FixedTable(nn.Module):           # SinusoidalPositionalEmbedding with a fixed n-D table (non-persistent buffer)
    def __init__(self, dim: int, shape):
        super().__init__()
        self.register_buffer("pos_emb", sincos_nd_table(dim, shape).unsqueeze(0), persistent=False)


class _AbsPosEmb(nn.Module):            # dit_base.py:504-525 (SinusoidalPositionalEmbedding, learnable or fixed)
    def __init__(self, dim: int, n_tokens: int, learnable: bool):
        super().__init__()
        if learnable:
            self.pos_emb = nn.Parameter(torch.zeros(1, n_tokens, dim).normal_(std=0.02))
        else:
            self.register_buffer("pos_emb", sincos_1d_table(dim, n_tokens).unsqueeze(0), persistent=False)


class _DiTBase(nn.Module):
    def __init__(self, dim: int, depth: int, spatial_mlp_ratio: Optional[float], out_channels: int,
                 pos_emb_type: str = "rope_3d", n_tokens: int = 0, factorized: bool = False,
                 mlp_ratio: Optional[float] = 4.0, grid=(1, 1), max_frames: int = 1, matrix: Optional[dict] = None):
        super().__init__()
        if matrix is not None:                             # dit_base.py:254-258 (sinusoidal_2d), :159-222
            self.pos_emb = _FixedTable(dim, tuple(grid))
            if matrix["block"] == "matrix":
                mk = lambda: _MatrixBlock(grid[0] * grid[1], dim, matrix["embed_col_dim"], mlp_ratio, matrix["use_bias"],
                                          matrix["fixed_u"])
            else:                                          # dit_base.py:27-31 `matrix_blocks`
                mk = lambda: _MatrixTokenBlock(grid[0] * grid[1], dim, matrix["embed_col_dim"], mlp_ratio,
                                               matrix["block"] == "matrix_cross")
            if matrix["full"]:
                self.blocks = nn.ModuleList([mk() for _ in range(depth)])
            else:
                self.blocks = nn.ModuleList([_Block(dim, spatial_mlp_ratio) for _ in range(depth)])
                self.temporal_blocks = nn.ModuleList([mk() for _ in range(depth)])
            self.final_layer = _FinalLayer(dim, out_channels)
            return
        # dit_base.py:156: the positional embedding is registered BEFORE the blocks (named_parameters() order)
        if pos_emb_type in ("learned_1d", "sinusoidal_1d"):
            self.pos_emb = _AbsPosEmb(dim, n_tokens, pos_emb_type == "learned_1d")
        elif pos_emb_type == "sinusoidal_factorized":      # dit_base.py:265-274
            self.spatial_pos_emb = _FixedTable(dim, tuple(grid))
            self.temporal_pos_emb = _FixedTable(dim, (max_frames,))
        # dit_base.py:185,192 — "full" and spatial blocks take spatial_mlp_ratio (None ⇒ no MLP; fork quirk Q2)
        self.blocks = nn.ModuleList([_Block(dim, spatial_mlp_ratio) for _ in range(depth)])
        if factorized:                                     # dit_base.py:196-222: temporal blocks take mlp_ratio, no RoPE
            self.temporal_blocks = nn.ModuleList([_Block(dim, mlp_ratio) for _ in range(depth)])
        self.final_layer = _FinalLayer(dim, out_channels)


def rope_axis_dims(head_dim: int):
    """embeddings.py:254-277: rotary widths of the (t, h, w) axes."""
    half = head_dim // 2
    q, r = divmod(half, 3)
    dims = {0: (q, q, q), 1: (q + 1, q, q), 2: (q, q + 1, q + 1)}[r]
    return tuple(2 * d for d in dims)


def rope_cos_sin_table(head_dim: int, sizes, theta: float = 10000.0) -> torch.Tensor:
    """[T*H*W, head_dim/2, 2] (cos, sin) of the pair angles (embeddings.py:156-213; both elements of an
    interleaved pair share one angle, so only every second column of the reference table is kept)."""
    T, H, W = sizes
    parts = []
    for axis, (dim, n) in enumerate(zip(rope_axis_dims(head_dim), sizes)):
        inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        ang = torch.arange(n, dtype=torch.float32)[:, None] * inv[None, :]
        shape = [1, 1, 1, dim // 2]
        shape[axis] = n
        parts.append(ang.reshape(shape).expand(T, H, W, dim // 2))
    ang = torch.cat(parts, dim=-1).reshape(T * H * W, head_dim // 2)
    return torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()


def rope_1d_cos_sin_table(dim: int, n: int, theta: float = 10000.0) -> torch.Tensor:
    """[n, dim/2, 2] (cos, sin) of RotaryEmbedding1D's pair angles (embeddings.py:218-231, 189-198)."""
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    ang = torch.arange(n, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class DiT3D(nn.Module):
    def __init__(self, cfg, x_shape, max_tokens: int, external_cond_type: Optional[str] = None,
                 external_cond_num_classes: Optional[int] = None, external_cond_dim: int = 0,
                 use_causal_mask: bool = True):
        if use_causal_mask:
            raise NotImplementedError("Causal masking is not yet implemented for DiT3D backbone")  # dit3d.py:23-26
        super().__init__()
        cfg = to_config(cfg)
        self.pos_emb_type = cfg.get("pos_emb_type", "rope_3d")
        self.variant = cfg.get("variant", "full")
        self.factorized = self.variant in ("factorized_encoder", "factorized_attention")   # one code path in the fork
        self.matrix = self.variant in ("full_matrix_attention", "factorized_matrix_attention")
        allowed = ("sinusoidal_2d",) if self.matrix else \
            ("learned_1d", "sinusoidal_1d", "sinusoidal_factorized") if self.factorized else \
            ("rope_3d", "learned_1d", "sinusoidal_1d")
        if not (self.variant == "full" or self.factorized or self.matrix) or self.pos_emb_type not in allowed:
            raise NotImplementedError(
                "dfot_b200 DiT3D supports variant=full with pos_emb_type rope_3d (the default dit3d.yaml), learned_1d or "
                "sinusoidal_1d, the factorized_encoder / factorized_attention variants with sinusoidal_factorized "
                "(dit3d_factorized_attention.yaml), learned_1d or sinusoidal_1d, and the matrix-attention variants with "
                "sinusoidal_2d (dit3d_full_matrix.yaml, dit3d_factorized_matrix.yaml); rope with a factorized variant and "
                "sinusoidal_3d (both assert in the fork itself) are not built")
        if self.matrix:
            self._check_matrix_cfg(cfg)
        self.cfg = cfg
        self.x_shape = list(x_shape)
        self.max_tokens = max_tokens
        self.external_cond_type = external_cond_type
        self.external_cond_num_classes = external_cond_num_classes
        self.external_cond_dim = external_cond_dim or 0
        self.use_causal_mask = use_causal_mask
        self.patch_size = cfg.patch_size
        C, H, W = self.x_shape
        self.num_patches_h, self.num_patches_w = H // self.patch_size, W // self.patch_size
        self.num_patches = self.num_patches_h * self.num_patches_w
        self.hidden_size = D = cfg.embed_row_dim if self.matrix else cfg.hidden_size      # dit3d.py:113-118
        self.depth = cfg.depth
        self.has_token_attention = self.variant != "full_matrix_attention"     # plain DiT blocks somewhere in the stack
        if self.has_token_attention:
            self.num_heads = cfg.num_heads
            self.head_dim = D // self.num_heads
            assert D % self.num_heads == 0, "dim should be divisible by num_heads"      # dit_blocks.py:66
            if self.head_dim not in (64, 72, 128):
                raise NotImplementedError(f"head_dim {self.head_dim} unsupported by the attention kernel (64, 72, 128)")
        else:
            self.num_heads, self.head_dim = 0, 0
        self.external_cond_dropout = cfg.get("external_cond_dropout", 0.0)

        self.noise_level_pos_embedding = _NoiseLevelEmbedding(256, D, bool(cfg.get("use_fourier_noise_embedding",
                                                                                  False)))
        self.label_cond = bool(self.external_cond_dim) and external_cond_type == "label"
        if self.label_cond:
            self.external_cond_embedding = _LabelEmbedding(external_cond_num_classes, D, self.external_cond_dropout)
        elif self.external_cond_dim:
            if external_cond_type != "action":
                raise ValueError(f"Unknown external condition type: {external_cond_type}. "
                                 "Supported types are 'label' and 'action'.")
            self.external_cond_embedding = (_TimestepMLP(self.external_cond_dim, D) if self.external_cond_dropout == 0
                                            else _CondEmbeddingDropout(self.external_cond_dim, D))
        else:
            self.external_cond_embedding = None
        self.patch_embedder = _PatchEmbed(C, D, self.patch_size)
        matrix = None
        if self.matrix:
            matrix = dict(full=self.variant == "full_matrix_attention", embed_col_dim=self.matrix_cols,
                          use_bias=bool(cfg.use_bias), fixed_u=cfg.get("fixed_u", None), block=self.matrix_block)
        self.dit_base = _DiTBase(D, self.depth, cfg.get("spatial_mlp_ratio", None), self.patch_size ** 2 * C,
                                 self.pos_emb_type, max_tokens * self.num_patches, factorized=self.factorized,
                                 mlp_ratio=cfg.get("mlp_ratio", 4.0), grid=(self.num_patches_h, self.num_patches_w),
                                 max_frames=max_tokens, matrix=matrix)
        self.use_rope = self.pos_emb_type == "rope_3d"
        self.use_mlp = self.dit_base.blocks[0].use_mlp
        if (self.factorized or self.matrix) and max_tokens > 128:
            raise NotImplementedError("factorized / matrix DiT3D: temporal attention over more than 128 frames is not built")
        self._init_embedders()
        self._packed = None
        self._packed_key = None
        self._ws = {}
        # One CUDA graph per (rows, frames, dtypes, conditioning) signature: a forward is ~400 launches of 5-300 us
        # kernels, so replaying a captured graph removes the host launch cost from the sampling loop.
        self.use_cuda_graph = True
        self._graphs = {}
        # K1 writes only the bf16 copy of the modulated tokens plus the rows' (mean, rstd); the gated-residual GEMM epilogue
        # rebuilds the fp32 residual base from x in place (118 -> 71 MB per K1 launch at K600).  DFOT_DIT_REBUILD_BASE=0: the
        # round-1 path (K1 stores the fp32 copy) — same bits.
        self.rebuild_residual_base = os.environ.get("DFOT_DIT_REBUILD_BASE", "1") != "0"

    def _check_matrix_cfg(self, cfg) -> None:
        """dit_base.py:129-149 (the reference's own assertions) + what the kernels cover.  With one row per column head
        (embed_col_dim == num_col_heads — every shipped matrix configuration has both = 1) a head's feature is one
        [1, head_row_dim] row and `flatten_matrix_rope` / `matrix_multi_token` do not change the computation (checked
        against the executed reference, oracle/make_goldens_matrix.py).  With n = embed_col_dim / num_col_heads > 1 rows per
        column head (dit_blocks.py:312-340): `matrix_multi_token` makes every one of the embed_col_dim rows its own sequence
        (scale head_row_dim^-1/2) — the same kernel sequence as n = 1; otherwise a head's feature is the flattened
        [n, head_row_dim] block (scale (n * head_row_dim)^-1/2), reached by regrouping the frame-level q | k | v rows
        (_matrix_core), rotated per row or — `flatten_matrix_rope` — as one n * head_row_dim wide vector."""
        self.matrix_block = cfg.get("matrix_block")
        assert self.matrix_block in ("matrix", "matrix_self", "matrix_cross"), f"Unknown matrix block {self.matrix_block}"
        for k in ("embed_col_dim", "embed_row_dim", "num_col_heads", "num_row_heads", "spatial_mlp_ratio", "use_bias"):
            assert cfg.get(k) is not None, f"{k} must be specified for matrix attention"
        assert cfg.embed_row_dim % cfg.num_row_heads == 0, "embed_row_dim must be divisible by num_row_heads"
        assert cfg.embed_col_dim % cfg.num_col_heads == 0, "embed_col_dim must be divisible by num_col_heads"
        if cfg.get("flatten_matrix_rope") and cfg.get("matrix_multi_token"):
            raise AssertionError("flatten_rope and multi_token cannot be used together.")      # dit_blocks.py:253
        self.matrix_cols = cfg.embed_col_dim
        self.matrix_heads = cfg.num_row_heads
        self.matrix_head_dim = cfg.embed_row_dim // cfg.num_row_heads
        self.matrix_rope = bool(cfg.get("use_temporal_rope", False))
        n = cfg.embed_col_dim // cfg.num_col_heads
        # rows of a column head that one attention head's feature spans (1: every row is a sequence of its own)
        self.matrix_group = 1 if cfg.get("matrix_multi_token") else n
        self.matrix_flatten_rope = bool(cfg.get("flatten_matrix_rope")) and self.matrix_group > 1
        self.matrix_feature_dim = self.matrix_group * self.matrix_head_dim      # the attention kernel's head_dim
        if self.matrix_feature_dim not in (64, 72, 128) or self.matrix_head_dim % 8:
            raise NotImplementedError(f"matrix attention feature width {self.matrix_feature_dim} (rows per head x "
                                      f"head_row_dim {self.matrix_head_dim}) unsupported by the attention kernel "
                                      "(64, 72, 128)")
        if self.matrix_block != "matrix" and self.matrix_head_dim not in (64, 72, 128):
            # attn2 of a MatrixSelf / MatrixCrossDiTBlock: num_row_heads heads of embed_row_dim / num_row_heads inside a frame
            raise NotImplementedError(f"matrix_block={self.matrix_block}: token attention head dim {self.matrix_head_dim} "
                                      "unsupported by the attention kernel (64, 72, 128)")
        # qkv_bias [embed_col_dim, 3E]: one bias row per column row.  One column row: the GEMM's bias vector; more: Mc extra
        # one-hot input columns select the row's bias out of Mc extra weight columns (_pack_matrix_block)
        self.matrix_bias_cols = _pad8(self.matrix_cols) if (cfg.use_bias and self.matrix_block == "matrix"
                                                            and self.matrix_cols > 1) else 0

    # dit3d.py:91-108
    def _init_embedders(self):
        w = self.patch_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        nn.init.zeros_(self.patch_embedder.proj.bias)
        mlps = [self.noise_level_pos_embedding]
        if self.external_cond_embedding is not None:
            mlps.append(self.external_cond_embedding)
        for root in mlps:
            for m in root.modules():
                if isinstance(m, nn.Linear):
                    nn.init.normal_(m.weight, std=0.02)
                    nn.init.zeros_(m.bias)

    @property
    def n_tokens_per_frame(self) -> int:
        return self.num_patches

    def _ordered_blocks(self):
        """Blocks in execution order with their kind: "full" (all tokens of a row), or per layer "spatial" then "temporal"."""
        mkind = self.matrix_block if self.matrix else None     # "matrix" | "matrix_self" | "matrix_cross"
        if self.variant == "full_matrix_attention":
            return [(mkind, b) for b in self.dit_base.blocks]
        if not (self.factorized or self.matrix):
            return [("full", b) for b in self.dit_base.blocks]
        out = []
        for sb, tb in zip(self.dit_base.blocks, self.dit_base.temporal_blocks):
            out += [("spatial", sb), (mkind if self.matrix else "temporal", tb)]
        return out

    # ------------------------------------------------------------------ weight packing
    def _version_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + \
               tuple((b.data_ptr(), b._version) for b in self.buffers())

    def packed(self):
        key = self._version_key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        dev = self.patch_embedder.proj.weight.device
        ops.require_cuda(dev, "DiT3D")                          # no CPU implementation exists
        D, C, p = self.hidden_size, self.x_shape[0], self.patch_size
        bf = lambda w: ops.cast_bf16(w.detach().float().contiguous())
        f32 = lambda b: b.detach().float().contiguous()
        P = {}
        te = self.noise_level_pos_embedding.embedding
        P["t1_w"], P["t1_b"], P["t2_w"], P["t2_b"] = bf(te.linear_1.weight), f32(te.linear_1.bias), \
            bf(te.linear_2.weight), f32(te.linear_2.bias)
        if hasattr(self.noise_level_pos_embedding, "timesteps"):
            P["four_f"] = f32(self.noise_level_pos_embedding.timesteps.freqs)
            P["four_p"] = f32(self.noise_level_pos_embedding.timesteps.phases)
        if self.label_cond:
            P["label_table"] = f32(self.external_cond_embedding.embedding_table.weight)
        elif self.external_cond_embedding is not None:
            ce = self.external_cond_embedding if self.external_cond_dropout == 0 else self.external_cond_embedding.embedding
            kc = _pad8(self.external_cond_dim)
            w1 = torch.zeros((D, kc), device=dev)
            w1[:, : self.external_cond_dim] = ce.linear_1.weight.detach().float()
            P["c1_w"], P["c1_b"], P["c2_w"], P["c2_b"] = bf(w1), f32(ce.linear_1.bias), bf(ce.linear_2.weight), \
                f32(ce.linear_2.bias)
        kp = _pad8(C * p * p)
        wp = torch.zeros((D, kp), device=dev)
        wp[:, : C * p * p] = self.patch_embedder.proj.weight.detach().float().reshape(D, -1)
        P["pe_w"], P["pe_b"] = bf(wp), f32(self.patch_embedder.proj.bias)
        mods_w, mods_b = [], []
        for _, blk in self._ordered_blocks():
            for norm in blk.adaln_norms():
                mods_w.append(norm.modulation[-1].weight)
                mods_b.append(norm.modulation[-1].bias)
        fl = self.dit_base.final_layer
        mods_w.append(fl.norm_final.modulation[-1].weight)
        mods_b.append(fl.norm_final.modulation[-1].bias)
        P["mod_w"] = bf(torch.cat([w.detach().float() for w in mods_w], 0))
        P["mod_b"] = torch.cat([b.detach().float() for b in mods_b], 0).contiguous()
        P["blocks"] = []
        if not self.use_rope:
            # absolute position table: the residual operand of the patch-embed GEMM; without RoPE the QKV epilogue is a plain
            # bf16 store, so the softmax scale (x log2 e: the attention kernel exponentiates in base 2) is folded into W_q, b_q
            if self.pos_emb_type == "sinusoidal_factorized":   # spatial table per frame now, temporal table before the
                P["pos"] = f32(self.dit_base.spatial_pos_emb.pos_emb[0]).repeat(self.max_tokens, 1)   # first temporal block
                P["tpos"] = f32(self.dit_base.temporal_pos_emb.pos_emb[0])
            elif self.pos_emb_type == "sinusoidal_2d":         # dit_base.py:356-362: the spatial table, per frame
                P["pos"] = f32(self.dit_base.pos_emb.pos_emb[0]).repeat(self.max_tokens, 1)
            else:
                P["pos"] = f32(self.dit_base.pos_emb.pos_emb[0])
            qs = torch.ones((3 * D, 1), device=dev)
            qs[:D] = LOG2E / math.sqrt(max(self.head_dim, 1))
        for kind, blk in self._ordered_blocks():
            if kind.startswith("matrix"):
                P["blocks"].append(self._pack_matrix_block(blk, kind, bf, f32, dev))
                continue
            qw, qb = blk.attn.qkv.weight.detach().float(), blk.attn.qkv.bias.detach().float()
            if not self.use_rope:
                qw, qb = qw * qs, qb * qs[:, 0]
            d = dict(kind=kind, n_norms=len(blk.adaln_norms()), qkv_w=bf(qw), qkv_b=f32(qb), proj_w=bf(blk.attn.proj.weight),
                     proj_b=f32(blk.attn.proj.bias))
            if blk.use_mlp:
                d.update(fc1_w=bf(blk.mlp.fc1.weight), fc1_b=f32(blk.mlp.fc1.bias), fc2_w=bf(blk.mlp.fc2.weight),
                         fc2_b=f32(blk.mlp.fc2.bias))
            P["blocks"].append(d)
        no = _pad8(p * p * C)
        wf = torch.zeros((no, D), device=dev)
        wf[: p * p * C] = fl.linear.weight.detach().float()
        bfin = torch.zeros((no,), device=dev)
        bfin[: p * p * C] = fl.linear.bias.detach().float()
        P["fin_w"], P["fin_b"] = bf(wf), bfin
        if self.use_rope:
            P["rope"] = rope_cos_sin_table(self.head_dim, (self.max_tokens, self.num_patches_h, self.num_patches_w)).to(dev)
        if self.matrix and self.matrix_rope:
            if not self.matrix_flatten_rope:         # (the flattened table depends on T: built with the workspace)
                P["mrope"] = rope_1d_cos_sin_table(self.matrix_head_dim, self.max_tokens).to(dev)
        self._packed, self._packed_key = P, key
        return P

    def _pack_matrix_block(self, blk, kind, bf, f32, dev):
        """MatrixDiTBlock weights in kernel layout: the v factors as [N, K] bf16 GEMM weights (W = v^T), the u factors as
        f32 tables of the two patch kernels.  Without the temporal RoPE the QKV epilogue is a plain bf16 store, so the
        softmax scale (x log2 e) is folded into the q columns (as for the absolute-position DiT blocks).  The token
        attention of a matrix_self / matrix_cross block (attn2, never rotated) gets the same folding; a cross block's
        q_proj / kv_proj stay two GEMMs (different inputs) that write the column slabs of one [M, 3D] q|k|v buffer."""
        a, E, Mc, Pn = blk.attn if kind == "matrix" else blk.attn1, self.hidden_size, self.matrix_cols, self.num_patches
        qw = a.qkv_v.detach().float().t().contiguous()                 # [3E, D]
        qb = torch.zeros((3 * E,), device=dev)
        if hasattr(a, "qkv_bias") and self.matrix_bias_cols:      # [Mc, 3E] -> weight columns D .. D+Mc (bf16 like the weights)
            qw = torch.cat([qw, a.qkv_bias.detach().float().t(),
                            torch.zeros((3 * E, self.matrix_bias_cols - Mc), device=dev)], dim=1)
        elif hasattr(a, "qkv_bias"):
            qb = a.qkv_bias.detach().float()[0].clone()
        if not self.matrix_rope:
            scale = LOG2E / math.sqrt(self.matrix_feature_dim)
            qw[:E] *= scale
            qb[:E] *= scale
        eye = torch.eye(Pn, device=dev) if a.fixed_u == "identity" else None
        d = dict(kind=kind, n_norms=len(blk.adaln_norms()), qkv_w=bf(qw), qkv_b=f32(qb),
                 proj_w=bf(a.proj_v.detach().float().t().contiguous()),
                 qkv_u=f32(eye if eye is not None else a.qkv_u).reshape(Pn, Mc).contiguous(),
                 proj_u=f32(eye if eye is not None else a.proj_u).reshape(Mc, Pn).contiguous(),
                 proj_bias=f32(a.proj_bias) if hasattr(a, "proj_bias") else None)
        if kind != "matrix":
            t, ts = blk.attn2, LOG2E / math.sqrt(self.matrix_head_dim)
            if kind == "matrix_cross":
                d.update(q2_w=bf(t.q_proj.weight.detach().float() * ts), q2_b=f32(t.q_proj.bias.detach().float() * ts),
                         kv2_w=bf(t.kv_proj.weight), kv2_b=f32(t.kv_proj.bias))
            else:
                qs2 = torch.ones((3 * E, 1), device=dev)
                qs2[:E] = ts
                d.update(qkv2_w=bf(t.qkv.weight.detach().float() * qs2), qkv2_b=f32(t.qkv.bias.detach().float() * qs2[:, 0]))
            d.update(proj2_w=bf(t.proj.weight), proj2_b=f32(t.proj.bias))
        if blk.use_mlp:
            d.update(fc1_w=bf(blk.mlp.fc1.weight), fc1_b=f32(blk.mlp.fc1.bias), fc2_w=bf(blk.mlp.fc2.weight),
                     fc2_b=f32(blk.mlp.fc2.bias))
        return d

    def _workspace(self, R: int, T: int, dev, out_dtype):
        key = (R, T, str(dev), out_dtype)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        D, C, p = self.hidden_size, self.x_shape[0], self.patch_size
        M, RT = R * T * self.num_patches, R * T
        blocks = self._ordered_blocks()
        n_mod = sum(3 * len(b.adaln_norms()) for _, b in blocks) + 2
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        bf, f32 = torch.bfloat16, torch.float32
        ws = dict(feat=e((RT, 256), bf), e1=e((RT, D), bf), emb=e((RT, D), f32), cact=e((RT, D), bf),
                  mod=e((RT, n_mod * D), f32), patches=torch.zeros((M, _pad8(C * p * p)), dtype=bf, device=dev),
                  x=e((M, D), f32), y=e((M, D), f32), y16=e((M, D), bf), qkv=e((M, 3 * D), bf), att=e((M, D), bf),
                  tok=e((M, _pad8(p * p * C)), f32), out=e((R, T, *self.x_shape), out_dtype), stats=e((M, 2), f32))
        hidden = max([b.mlp.fc1.out_features for _, b in blocks if b.use_mlp], default=0)
        if hidden:
            ws["h"] = e((M, hidden), bf)
        if self.factorized:
            # temporal blocks: the token stream in (row, patch, frame) order, the block's modulation columns per token and the
            # frame index of every token of that order (m = (r*P + p)*T + t  ->  frame r*T + t)
            Pn = self.num_patches
            m = torch.arange(M, device=dev)
            ncol_t = (6 if self.dit_base.temporal_blocks[0].use_mlp else 3) * D
            ws.update(xt=e((M, D), f32), mod_tok=e((M, ncol_t), f32), frame_of_tok=(m // (Pn * T)) * T + m % T)
        if self.matrix:      # frame-level rows (row, column head, frame) of the matrix attention
            Mf = R * self.matrix_cols * T
            ws.update(ms=e((Mf, D), bf), mqkv=e((Mf, 3 * D), bf), matt=e((Mf, D), bf), mz=e((Mf, D), f32))
            if self.matrix_bias_cols:                      # [u^T y | one-hot(column row)]: the QKV GEMM's input with bias columns
                msb = torch.zeros((Mf, D + self.matrix_bias_cols), dtype=bf, device=dev)
                hot = msb.view(R, self.matrix_cols, T, -1)
                for mc in range(self.matrix_cols):
                    hot[:, mc, :, D + mc] = 1.0
                ws["ms_b"] = msb
            if self.matrix_group > 1:                      # q | k | v and the attention output regrouped per column head
                ws.update(mqkv_g=e((Mf // self.matrix_group, 3 * D * self.matrix_group), bf),
                          matt_g=e((Mf // self.matrix_group, D * self.matrix_group), bf))
            if self.matrix_flatten_rope:
                # dit_blocks.py:316-318: the flattened [g, d] feature of a head rotates as one g*d wide vector, i.e. row j of
                # a column head takes the pair angles j*d/2 .. (j+1)*d/2 of the wider table: one table row per (column row,
                # frame), addressed by the QKV epilogue as row % (Mc * T)
                g, d2 = self.matrix_group, self.matrix_head_dim // 2
                t = rope_1d_cos_sin_table(self.matrix_feature_dim, T).view(T, g, d2, 2)
                ws["mrope"] = t.permute(1, 0, 2, 3).repeat(self.matrix_cols // g, 1, 1, 1).reshape(-1, d2, 2).contiguous().to(dev)
            if self.matrix_block == "matrix_cross":        # attn1's output, the k / v source of attn2
                ws.update(x1=e((M, D), f32), x1_16=e((M, D), bf))
        if not self.use_rope:
            ws["pos_rows"], ws["pos_key"] = e((M, D), f32), None     # the table repeated per row (filled lazily)
        if self.external_cond_embedding is not None:
            ws.update(cin=torch.zeros((RT, _pad8(self.external_cond_dim)), dtype=bf, device=dev), c1=e((RT, D), bf),
                      cemb=e((RT, D), f32))
        self._ws[key] = ws
        return ws

    def _fill_pos_rows(self, ws, R: int, Ntok: int) -> None:
        """The absolute position table repeated for every row of the batch (the GEMM epilogue's residual operand has no
        modulo addressing).  Re-filled when the weights change; never inside a graph capture (forward() calls it first)."""
        if ws["pos_key"] == self._packed_key:
            return
        if ws["pos_rows"].is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("DiT3D: the position rows must be filled before a CUDA-graph capture")
        D = self.hidden_size
        ws["pos_rows"].view(R, Ntok, D).copy_(self.packed()["pos"][:Ntok].unsqueeze(0).expand(R, Ntok, D))
        ws["pos_key"] = self._packed_key

    # ------------------------------------------------------------------ forward
    def input_buffer(self, R: int, T: int, dtype, device) -> torch.Tensor:
        """Static input tensor of the captured graph for this signature; writing the branch inputs straight into it
        (as the fused sampler kernel does) avoids a copy per step."""
        key = ("in", R, T, dtype, str(device))
        buf = self._ws.get(key)
        if buf is None:
            buf = torch.empty((R, T, *self.x_shape), dtype=dtype, device=device)
            self._ws[key] = buf
        return buf

    @torch.no_grad()
    def forward(self, x: torch.Tensor, noise_levels: torch.Tensor, external_cond: Optional[torch.Tensor] = None,
                external_cond_mask: Optional[torch.Tensor] = None, out_dtype=torch.float32) -> torch.Tensor:
        """x [R,T,C,H,W] f32|bf16; noise_levels [R,T] int64 (discrete) or f32 (continuous: precond*logsnr).
        Returns a tensor shaped like x (a workspace buffer that the next call overwrites)."""
        ops.require_cuda(x.device, "DiT3D.forward")
        if not self.use_cuda_graph or torch.cuda.is_current_stream_capturing():
            return self._forward_impl(x, noise_levels, external_cond, external_cond_mask, out_dtype)
        R, T = x.shape[:2]
        levels = noise_levels if noise_levels.dtype in (torch.int64, torch.float32) else noise_levels.float()
        use_mask = external_cond_mask is not None and external_cond is not None and self.external_cond_dropout != 0
        sig = (R, T, x.dtype, levels.dtype, external_cond is not None, use_mask, out_dtype, str(x.device),
               ops.latency_mode())
        st = self._graphs.get(sig)
        if st is None:   # first call with this signature: eager (also warms up lazy kernel attributes)
            self._graphs[sig] = {"graph": None, "key": self._version_key()}
            return self._forward_impl(x, noise_levels, external_cond, external_cond_mask, out_dtype)
        if st["key"] != self._version_key():   # weights changed: drop the stale graph
            st["graph"], st["key"] = None, self._version_key()
        xin = self.input_buffer(R, T, x.dtype, x.device)
        if x.data_ptr() != xin.data_ptr():
            xin.copy_(x)
        if not self.use_rope:
            self.packed()
            self._fill_pos_rows(self._workspace(R, T, x.device, out_dtype), R, T * self.num_patches)
        if st["graph"] is None:
            self.packed()
            st["levels"] = torch.empty_like(levels)
            st["cond"] = None if external_cond is None else torch.empty_like(external_cond, dtype=torch.float32)
            st["mask"] = torch.empty_like(external_cond_mask) if use_mask else None
        st["levels"].copy_(levels)
        # conditions / masks are constant over the steps of a window: refresh the graph's static copies only when the source
        # tensor (or its contents: _version) changed
        for name, src in (("cond", external_cond), ("mask", external_cond_mask)):
            if st[name] is not None:
                key = (src.data_ptr(), src._version, tuple(src.shape))
                if st.get(name + "_key") != key:
                    st[name].copy_(src)
                    st[name + "_key"] = key
        if st["cond"] is not None:
            # the condition embedding is constant over the sampling steps of a window: it is computed here, outside the
            # graph, when the condition tensor (or the weights) changed — keyed on the workspace, which the graphs of
            # different signatures share; the source tensor is kept alive so that its address cannot be re-used
            Pk, ws = self.packed(), self._workspace(R, T, x.device, out_dtype)
            ckey = (external_cond.data_ptr(), external_cond._version, tuple(external_cond.shape), self._packed_key)
            if ws.get("cemb_key") != ckey:
                self._cond_embedding(Pk, ws, st["cond"], R, T)
                ws["cemb_key"], ws["cemb_src"] = ckey, external_cond
        if st["graph"] is None:
            g = torch.cuda.CUDAGraph()
            n0 = _abi.launch_count()
            with torch.cuda.graph(g):
                st["out"] = self._forward_impl(xin, st["levels"], st["cond"], st["mask"], out_dtype, cond_cached=True)
            st["graph"], st["kernels"] = g, _abi.launch_count() - n0
        st["graph"].replay()
        ops.count_replayed_launches(st["kernels"])
        return st["out"]

    def _cond_embedding(self, Pk, ws, external_cond: torch.Tensor, R: int, T: int) -> None:
        """ws["cemb"] <- the per-frame embedding of the external condition (before the mask).  It does not depend on the
        noise levels, so under a CUDA graph it is computed once per condition tensor, outside the graph (forward())."""
        ws["cemb_key"] = None            # (whoever caches the result sets the key afterwards)
        if self.label_cond:
            # dit3d.py:171-173: emb + table[labels]; labels [R, 1] (one per clip, ucf_101.py:304-309) or [R, T]; the
            # mask is not passed to the label embedding.  The row gather is data movement (torch indexing).
            lab = external_cond.reshape(R, -1).long()
            ws["cemb"].view(R, T, -1).copy_(Pk["label_table"][lab].expand(R, T, -1))
        else:
            ws["cin"][:, : self.external_cond_dim] = external_cond.reshape(R * T, -1).to(torch.bfloat16)
            ops.gemm_bf16(ws["cin"], Pk["c1_w"], ws["c1"], ops.EPI_SILU_BF16, bias=Pk["c1_b"])
            ops.gemm_bf16(ws["c1"], Pk["c2_w"], ws["cemb"], ops.EPI_F32, bias=Pk["c2_b"])

    @torch.no_grad()
    def _forward_impl(self, x: torch.Tensor, noise_levels: torch.Tensor, external_cond: Optional[torch.Tensor] = None,
                      external_cond_mask: Optional[torch.Tensor] = None, out_dtype=torch.float32,
                      cond_cached: bool = False) -> torch.Tensor:
        R, T = x.shape[:2]
        C, H, W = self.x_shape
        D, p, Pn = self.hidden_size, self.patch_size, self.num_patches
        if T > self.max_tokens:
            raise ValueError(f"Input sequence length {T * Pn} exceeds the maximum length {self.max_tokens * Pn}")
        Pk = self.packed()
        ws = self._workspace(R, T, x.device, out_dtype)
        M, RT, Ntok = R * T * Pn, R * T, T * Pn
        x = x.contiguous()
        levels = noise_levels.contiguous()
        if levels.dtype not in (torch.int64, torch.float32):
            levels = levels.float()

        # --- tokens
        ops.patchify_bf16(x, ws["patches"], RT, C, H, W, p)
        if self.use_rope:
            ops.gemm_bf16(ws["patches"], Pk["pe_w"], ws["x"], ops.EPI_F32, bias=Pk["pe_b"])
        else:       # dit_base.py:352-353: x + pos_emb[:, :seq_len], as the residual operand of the patch-embed GEMM
            self._fill_pos_rows(ws, R, Ntok)
            ops.gemm_bf16(ws["patches"], Pk["pe_w"], ws["x"], ops.EPI_RESID_F32, bias=Pk["pe_b"], resid=ws["pos_rows"])
        # --- per-frame conditioning vector c = silu(noise_emb [+ cond_emb])
        ops.noise_features(levels, ws["feat"], Pk.get("four_f"), Pk.get("four_p"))
        ops.gemm_bf16(ws["feat"], Pk["t1_w"], ws["e1"], ops.EPI_SILU_BF16, bias=Pk["t1_b"])
        ops.gemm_bf16(ws["e1"], Pk["t2_w"], ws["emb"], ops.EPI_F32, bias=Pk["t2_b"])
        cemb, row_mask = None, None
        if external_cond is not None:
            if self.external_cond_embedding is None:
                raise ValueError("external_cond given but the backbone was built with external_cond_dim=0")
            cemb = ws["cemb"]
            if not cond_cached:
                self._cond_embedding(Pk, ws, external_cond, R, T)
            # embeddings.py:364-387: with dropout_prob == 0 the embedding is a plain MLP and ignores the mask; the label
            # embedding never sees it (dit3d.py:171-173)
            if not self.label_cond and external_cond_mask is not None and self.external_cond_dropout != 0:
                row_mask = external_cond_mask.to(torch.uint8).contiguous()
        ops.silu_sum_bf16(ws["emb"], cemb, row_mask, T, ws["cact"])
        ops.gemm_bf16(ws["cact"], Pk["mod_w"], ws["mod"], ops.EPI_F32, bias=Pk["mod_b"])
        mod, ldm = ws["mod"], ws["mod"].shape[1]

        # --- blocks
        q_scale = LOG2E / math.sqrt(max(self.head_dim, 1))
        if self._use_splitk(M):
            self._blocks_splitk(Pk, ws, mod, R, T, q_scale)
            ops.gemm_bf16(ws["y16"], Pk["fin_w"], ws["tok"], ops.EPI_F32, bias=Pk["fin_b"])
            ops.unpatchify(ws["tok"], ws["out"], RT, C, H, W, p)
            return ws["out"]
        col = 0
        xa, xb = ws["x"], ws["y"]
        first_temporal = True
        for bw in Pk["blocks"]:
            kind, ncol = bw["kind"], 3 * bw["n_norms"] * D
            xs, bmod, bld, bcol, tpf, n_seq, seq_len = xa, mod, ldm, col, Pn, R, Ntok
            if kind == "spatial":            # attention inside a frame: R*T sequences of P tokens, same token order
                n_seq, seq_len = RT, Pn
            elif kind == "temporal":         # dit_base.py:403-410: "(b t) p c -> (b p) t c", block, and back
                xs = ws["xt"]
                xs.view(R, Pn, T, D).copy_(xa.view(R, T, Pn, D).permute(0, 2, 1, 3))
                if first_temporal and "tpos" in Pk:      # temporal table, once, before the first temporal block
                    xs.view(R * Pn, T, D).add_(Pk["tpos"][:T])
                first_temporal = False
                torch.index_select(mod[:, col: col + ncol], 0, ws["frame_of_tok"], out=ws["mod_tok"])
                bmod, bld, bcol, tpf, n_seq, seq_len = ws["mod_tok"], ws["mod_tok"].shape[1], 0, 1, R * Pn, T
            if kind == "matrix":
                self._matrix_attention(bw, Pk, ws, xs, xb, mod, ldm, col, R, T)
            elif kind == "matrix_self":
                self._matrix_attention(bw, Pk, ws, xs, xb, mod, ldm, col, R, T)
                bcol += 3 * D
                self._frame_self_attention(bw, ws, xs, xb, mod, ldm, bcol, RT)
            elif kind == "matrix_cross":
                self._matrix_cross_attention(bw, Pk, ws, xs, xb, mod, ldm, col, R, T)
            else:
                self._token_attention(bw, Pk, ws, xs, xb, bmod, bld, bcol, tpf, n_seq, seq_len, Ntok, q_scale)
            if "fc1_w" in bw:
                bcol += 3 * D
                hbuf = ws["h"][:, : bw["fc1_w"].shape[0]]
                if self.rebuild_residual_base:
                    ops.adaln_layernorm(xs, bmod, bcol, bcol + D, tpf, y_bf16=ws["y16"], stats=ws["stats"])
                    ops.gemm_bf16(ws["y16"], bw["fc1_w"], hbuf, ops.EPI_GELU_BF16, bias=bw["fc1_b"])
                    self._gate_lnresid_gemm(hbuf, bw["fc2_w"], bw["fc2_b"], xs, ws["stats"], bmod, bld, bcol, tpf)
                else:
                    ops.adaln_layernorm(xs, bmod, bcol, bcol + D, tpf, y_f32=xb, y_bf16=ws["y16"])
                    ops.gemm_bf16(ws["y16"], bw["fc1_w"], hbuf, ops.EPI_GELU_BF16, bias=bw["fc1_b"])
                    ops.gemm_bf16(hbuf, bw["fc2_w"], xs, ops.EPI_GATE_RESID_F32, bias=bw["fc2_b"], resid=xb,
                                  gate=bmod[:, bcol + 2 * D:], ld_gate=bld, tokens_per_frame=tpf)
            col += ncol
            if kind == "temporal":
                xa.view(R, T, Pn, D).copy_(xs.view(R, Pn, T, D).permute(0, 2, 1, 3))
        # --- final layer + unpatchify
        ops.adaln_layernorm(xa, mod, col, col + D, Pn, y_bf16=ws["y16"])
        ops.gemm_bf16(ws["y16"], Pk["fin_w"], ws["tok"], ops.EPI_F32, bias=Pk["fin_b"])
        ops.unpatchify(ws["tok"], ws["out"], RT, C, H, W, p)
        return ws["out"]

    def _use_splitk(self, M: int) -> bool:
        """Latency mode (ops.set_latency_mode, off by default) in the latency regime (small-batch sampling: up to ~1k token
        rows, e.g. DMLab batch 1 = 256): the GEMMs that end a block half (proj, fc2: N = D columns — fewer tiles than SMs —
        and, for fc2, the longest k-loop of the block) deal their k-blocks over several CTAs and the gated residual moves
        into the AdaLN kernel that follows (ops.gemm_bf16_splitk / ops.splitk_gate_resid_adaln).  The number of splits
        depends on the row count, so a row's bits depend on its batch — which is why this is a mode and not the default."""
        return (ops.latency_mode() and self.variant == "full" and M <= ops.SPLITK_MAX_ROWS and self.hidden_size % 64 == 0
                and self.hidden_size <= 2048)

    def _blocks_splitk(self, Pk, ws, mod, R: int, T: int, q_scale: float) -> None:
        """The block loop of variant=full in the latency regime; leaves the final layer's modulated tokens in ws["y16"].
        Same arithmetic as the plain loop except for the summation order of the split k-loops.  The token stream x is never
        stored: its only reader is the next AdaLN (the residual base of a block half is the modulated tensor, quirk Q1), so
        the two fp32 buffers alternate as y (residual base) of consecutive halves."""
        D, Pn = self.hidden_size, self.num_patches
        Ntok, M = T * Pn, R * T * Pn
        y_cur, y_alt = ws["y"], ws["x"]
        if "parts" not in ws:                # split-K partial sums (at most 8 splits, ops.splitk_factor)
            ws["parts"] = torch.empty((M, 8 * D), dtype=torch.float32, device=ws["x"].device)
        parts, y16 = ws["parts"], ws["y16"]
        pending, col = None, 0               # (splits, bias, gate column) of the GEMM whose partial sums are in `parts`

        def norm(shift_col, want_f32=True):
            nonlocal y_cur, y_alt, pending
            if pending is None:              # first norm of the network: the patch-embed output is in ws["x"]
                ops.adaln_layernorm(ws["x"], mod, shift_col, shift_col + D, Pn, y_f32=y_cur if want_f32 else None, y_bf16=y16)
                return
            S, bias, gate_col = pending
            ops.splitk_gate_resid_adaln(parts, S, bias, y_cur, mod, gate_col, shift_col, shift_col + D, Pn,
                                        y_f32=y_alt if want_f32 else None, y_bf16=y16)
            y_cur, y_alt, pending = y_alt, y_cur, None

        for bw in Pk["blocks"]:
            norm(col)
            if self.use_rope:
                ops.gemm_bf16(y16, bw["qkv_w"], ws["qkv"], ops.EPI_QKV_ROPE_BF16, bias=bw["qkv_b"], rope_cs=Pk["rope"],
                              tokens_per_sample=Ntok, model_dim=D, head_dim=self.head_dim, q_scale=q_scale)
            else:
                ops.gemm_bf16(y16, bw["qkv_w"], ws["qkv"], ops.EPI_BF16, bias=bw["qkv_b"])
            ops.attention(ws["qkv"], ws["att"], R, Ntok, self.num_heads, self.head_dim)
            S = ops.splitk_factor(M, D, D)
            ops.gemm_bf16_splitk(ws["att"], bw["proj_w"], parts, S)
            pending = (S, bw["proj_b"], col + 2 * D)
            if "fc1_w" in bw:
                norm(col + 3 * D)
                hidden = bw["fc1_w"].shape[0]
                hbuf = ws["h"][:, :hidden]
                ops.gemm_bf16(y16, bw["fc1_w"], hbuf, ops.EPI_GELU_BF16, bias=bw["fc1_b"])
                S = ops.splitk_factor(M, D, hidden)
                ops.gemm_bf16_splitk(hbuf, bw["fc2_w"], parts, S)
                pending = (S, bw["fc2_b"], col + 5 * D)
                col += 6 * D
            else:
                col += 3 * D
        norm(col, want_f32=False)            # final layer's AdaLayerNorm (dit_blocks.py:533-542)

    def _token_attention(self, bw, Pk, ws, xs, xb, bmod, bld, bcol, tpf, n_seq, seq_len, Ntok, q_scale):
        """dit_blocks.py:488-507, first half of a DiTBlock: x <- y + gate * proj(attention(qkv(y))), y = modulate(LN(x))."""
        D = self.hidden_size
        rebuild = self.rebuild_residual_base
        if rebuild:
            ops.adaln_layernorm(xs, bmod, bcol, bcol + D, tpf, y_bf16=ws["y16"], stats=ws["stats"])
        else:
            ops.adaln_layernorm(xs, bmod, bcol, bcol + D, tpf, y_f32=xb, y_bf16=ws["y16"])
        if self.use_rope:
            ops.gemm_bf16(ws["y16"], bw["qkv_w"], ws["qkv"], ops.EPI_QKV_ROPE_BF16, bias=bw["qkv_b"],
                          rope_cs=Pk["rope"], tokens_per_sample=Ntok, model_dim=D, head_dim=self.head_dim,
                          q_scale=q_scale)
        else:
            ops.gemm_bf16(ws["y16"], bw["qkv_w"], ws["qkv"], ops.EPI_BF16, bias=bw["qkv_b"])
        ops.attention(ws["qkv"], ws["att"], n_seq, seq_len, self.num_heads, self.head_dim)
        # x1 = y + gate1 * proj(att)   (residual base is the modulated tensor — reference quirk Q1)
        if rebuild:
            self._gate_lnresid_gemm(ws["att"], bw["proj_w"], bw["proj_b"], xs, ws["stats"], bmod, bld, bcol, tpf)
        else:
            ops.gemm_bf16(ws["att"], bw["proj_w"], xs, ops.EPI_GATE_RESID_F32, bias=bw["proj_b"], resid=xb,
                          gate=bmod[:, bcol + 2 * D:], ld_gate=bld, tokens_per_frame=tpf)

    def _gate_lnresid_gemm(self, a, w, b, xs, stats, bmod, bld, bcol, tpf):
        """x <- modulate(LN(x)) + gate * (a w^T + b), in place: the residual base of a block half is the modulated tensor
        (quirk Q1), which K1 did not store in fp32 — the epilogue rebuilds it from x, the rows' (mean, rstd) and the frame's
        shift / scale vectors with K1's own expression (bit-identical to the stored copy: DFOT_EPI_GATE_LNRESID_F32)."""
        D = self.hidden_size
        ops.gemm_bf16(a, w, xs, ops.EPI_GATE_LNRESID_F32, bias=b, resid=xs, gate=bmod[:, bcol + 2 * D:], ld_gate=bld,
                      tokens_per_frame=tpf, ln_stats=stats, ln_shift=bmod[:, bcol:], ln_scale=bmod[:, bcol + D:])

    def _matrix_attention(self, bw, Pk, ws, xs, xb, mod, ldm, col, R: int, T: int):
        """dit_blocks.py:626-644 + 289-350, first half of a MatrixDiTBlock: x <- y + gate * (proj_u^T A(u^T y v) proj_v +
        proj_bias), attention A over the T frames of a row (see the module docstring)."""
        D, Pn, Mc = self.hidden_size, self.num_patches, self.matrix_cols
        ops.adaln_layernorm(xs, mod, col, col + D, Pn, y_f32=xb)
        self._matrix_core(bw, Pk, ws, xb, R, T)
        ops.patch_expand_gate_resid(xs, xb, ws["mz"], bw["proj_u"], bw["proj_bias"], mod[:, col + 2 * D:], ldm, R, T, Pn, Mc)

    def _frame_self_attention(self, bw, ws, xs, xb, mod, ldm, col, RT: int):
        """dit_blocks.py:874-877, second half of a MatrixSelfDiTBlock: x <- y + gate * proj(attention(qkv(y))) with the Pn
        tokens of a frame as one sequence, y = modulate(LN(x)) through norm2, no rotation."""
        D, Pn = self.hidden_size, self.num_patches
        ops.adaln_layernorm(xs, mod, col, col + D, Pn, y_f32=xb, y_bf16=ws["y16"])
        ops.gemm_bf16(ws["y16"], bw["qkv2_w"], ws["qkv"], ops.EPI_BF16, bias=bw["qkv2_b"])
        ops.attention(ws["qkv"], ws["att"], RT, Pn, self.matrix_heads, self.matrix_head_dim)
        ops.gemm_bf16(ws["att"], bw["proj2_w"], xs, ops.EPI_GATE_RESID_F32, bias=bw["proj2_b"], resid=xb,
                      gate=mod[:, col + 2 * D:], ld_gate=ldm, tokens_per_frame=Pn)

    def _matrix_cross_attention(self, bw, Pk, ws, xs, xb, mod, ldm, col, R: int, T: int):
        """dit_blocks.py:752-763, first half of a MatrixCrossDiTBlock: x1 = matrix attention of y (bare: no gate, no
        residual), then per frame x <- y + gate * proj(attention(q(y), k(x1), v(x1)))."""
        D, Pn, Mc = self.hidden_size, self.num_patches, self.matrix_cols
        ops.adaln_layernorm(xs, mod, col, col + D, Pn, y_f32=xb, y_bf16=ws["y16"])
        self._matrix_core(bw, Pk, ws, xb, R, T)
        ops.patch_expand_gate_resid(ws["x1"], None, ws["mz"], bw["proj_u"], bw["proj_bias"], None, 0, R, T, Pn, Mc)
        ops.cast_bf16(ws["x1"], ws["x1_16"])
        ops.gemm_bf16(ws["y16"], bw["q2_w"], ws["qkv"][:, :D], ops.EPI_BF16, bias=bw["q2_b"])
        ops.gemm_bf16(ws["x1_16"], bw["kv2_w"], ws["qkv"][:, D:], ops.EPI_BF16, bias=bw["kv2_b"])
        ops.attention(ws["qkv"], ws["att"], R * T, Pn, self.matrix_heads, self.matrix_head_dim)
        ops.gemm_bf16(ws["att"], bw["proj2_w"], xs, ops.EPI_GATE_RESID_F32, bias=bw["proj2_b"], resid=xb,
                      gate=mod[:, col + 2 * D:], ld_gate=ldm, tokens_per_frame=Pn)

    def _matrix_core(self, bw, Pk, ws, xb, R: int, T: int):
        """ws["mz"] <- A(u^T y v) proj_v for y = xb: MatrixAttention up to its `proj_u` factor (dit_blocks.py:289-344)."""
        D, Pn, Mc, g = self.hidden_size, self.num_patches, self.matrix_cols, self.matrix_group
        H, d = self.matrix_heads, self.matrix_head_dim
        ops.patch_mix_bf16(xb, bw["qkv_u"], ws["ms"], R, T, Pn, Mc)
        ms = ws["ms"]
        if self.matrix_bias_cols and bw["qkv_w"].shape[1] > D:
            ms = ws["ms_b"]
            ms[:, :D].copy_(ws["ms"])                      # (frame-level rows: R*Mc*T x D)
        if self.matrix_rope:
            flat = self.matrix_flatten_rope
            ops.gemm_bf16(ms, bw["qkv_w"], ws["mqkv"], ops.EPI_QKV_ROPE_BF16, bias=bw["qkv_b"],
                          rope_cs=ws["mrope"] if flat else Pk["mrope"], tokens_per_sample=Mc * T if flat else T,
                          model_dim=D, head_dim=d, q_scale=LOG2E / math.sqrt(self.matrix_feature_dim))
        else:
            ops.gemm_bf16(ms, bw["qkv_w"], ws["mqkv"], ops.EPI_BF16, bias=bw["qkv_b"])
        if g == 1:
            ops.attention(ws["mqkv"], ws["matt"], R * Mc, T, H, d)
        else:
            # dit_blocks.py:334-340: the g rows of a column head form one [g, d] feature per row head — regroup the frame-level
            # rows (row, column head, row j, frame) x (q|k|v, row head, d) into (row, column head, frame) x (q|k|v, row head,
            # j, d) (data movement on R*Mc*T rows), attend with g*d wide heads, and back
            C = Mc // g
            ws["mqkv_g"].view(R * C, T, 3, H, g, d).copy_(ws["mqkv"].view(R * C, g, T, 3, H, d).permute(0, 2, 3, 4, 1, 5))
            ops.attention(ws["mqkv_g"], ws["matt_g"], R * C, T, H, g * d)
            ws["matt"].view(R * C, g, T, H, d).copy_(ws["matt_g"].view(R * C, T, H, g, d).permute(0, 3, 1, 2, 4))
        ops.gemm_bf16(ws["matt"], bw["proj_w"], ws["mz"], ops.EPI_F32)
