"""UViT3DPose backbone (RE10K camera-pose-conditioned U-ViT) on hand-written sm_100a kernels.

Drop-in for the reference classes
    algorithms/dfot/backbones/u_vit/u_vit3d_pose.py:13-131 (UViT3DPose),
    algorithms/dfot/backbones/u_vit/u_vit3d.py:22-335      (UViT3D level structure, skips),
    algorithms/dfot/backbones/u_vit/u_vit_blocks.py        (EmbedInput, ProjectOutput, ResBlock, NormalizeWithCond,
                                                            TransformerBlock, Downsample, Upsample)
same constructor, same ``forward(x, noise_levels, external_cond, external_cond_mask)`` and the same
``state_dict()`` keys.  block_types ResBlock / TransformerBlock, pos_emb_type "rope" (the shipped configs).

B200-first execution model (not a translation):
  * every activation is channel-last ``[images, h, w, C] == [tokens, C]``: ResBlock levels and transformer levels share
    one layout, so the reference's NCHW <-> token rearranges (u_vit3d.py:199-235) do not exist;
  * 3x3 convolutions are implicit GEMMs on the tcgen05 kernel (4-D TMA boxes shifted per tap, zero fill = padding);
  * the FiLM modulation of every block is ``emb_layer(noise_emb[frame] + pose_emb[pixel])`` and emb_layer is linear, so
    it splits into a per-frame part (ONE small GEMM per forward for all blocks) and a per-pixel camera-pose part that is
    constant over the sampling steps of a window: it is computed once per window into an HBM-resident bf16 cache
    (~1 GB per conditioned sample at RE10K size; the reference recomputes ray encodings (377 MB/row) + PatchEmbed +
    every emb_layer at every step);
  * ray encodings are generated directly as PatchEmbed rows (``dfot_pose_ray_patches``), the [R,T,180,H,W] tensor of
    the reference (dfot_video_pose.py:64-110) is never materialised on the fast path.
"""
import math
import os
from typing import List, Optional

import torch
from torch import nn

from dfot_b200 import _abi, ops
from dfot_b200.config import to_config
from ..dit.dit3d import LOG2E, _Fourier, _NoiseLevelEmbedding, _PatchEmbed, _pad8, rope_cos_sin_table  # noqa: F401
# q/k-norm + RoPE ride on the QKV GEMM's epilogue from this width on (DFOT_UVIT_FUSE_QKNORM_MIN_CH pins it: benchmarking)
_FUSE_QKNORM_MIN_CH = int(os.environ.get("DFOT_UVIT_FUSE_QKNORM_MIN_CH", "1024"))


# ------------------------------------------------------------------ parameter containers (reference key names)
class _RMSNormW(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class _EmbedInput(nn.Module):
    def __init__(self, in_ch: int, dim: int, p: int):
        super().__init__()
        self.proj = nn.Conv2d(in_ch, dim, kernel_size=p, stride=p)


class _ProjectOutput(nn.Module):
    def __init__(self, dim: int, out_ch: int, p: int):
        super().__init__()
        self.proj = nn.ConvTranspose2d(dim, out_ch, kernel_size=p, stride=p)
        nn.init.zeros_(self.proj.weight)
        nn.init.zeros_(self.proj.bias)


class _ResBlock(nn.Module):                 # u_vit_blocks.py:56-94
    kind = "res"

    def __init__(self, ch: int, emb_dim: int):
        super().__init__()
        self.emb_layer = nn.Conv2d(emb_dim, ch * 2, kernel_size=(1, 1))
        self.in_layers = nn.Sequential(nn.GroupNorm(32, ch, eps=1e-6), nn.SiLU(),
                                       nn.Conv2d(ch, ch, kernel_size=(3, 3), padding=(1, 1)))
        self.out_norm = nn.GroupNorm(32, ch, eps=1e-6)
        self.out_rest = nn.Sequential(nn.SiLU(), nn.Conv2d(ch, ch, kernel_size=(3, 3), padding=(1, 1)))
        nn.init.zeros_(self.out_rest[1].weight)
        nn.init.zeros_(self.out_rest[1].bias)


class _NormalizeWithCond(nn.Module):        # u_vit_blocks.py:98-121
    def __init__(self, dim: int, emb_dim: int):
        super().__init__()
        self.emb_layer = nn.Linear(emb_dim, dim * 2)
        self.norm = _RMSNormW(dim)


class _TransformerBlock(nn.Module):         # u_vit_blocks.py:196-274
    kind = "transformer"

    def __init__(self, dim: int, heads: int, emb_dim: int):
        super().__init__()
        self.norm = _NormalizeWithCond(dim, emb_dim)
        self.fused_attn_mlp_proj = nn.Linear(dim, 7 * dim, bias=True)
        self.q_norm, self.k_norm = _RMSNormW(dim // heads), _RMSNormW(dim // heads)
        self.attn_out = nn.Linear(dim, dim, bias=True)
        self.mlp_out = nn.Sequential(nn.SiLU(), nn.Dropout(0.0), nn.Linear(4 * dim, dim, bias=True))
        for lin in (self.attn_out, self.mlp_out[2]):
            nn.init.zeros_(lin.weight)
            nn.init.zeros_(lin.bias)


class _Resample(nn.Module):                 # Downsample / Upsample: a 3x3 conv (u_vit_blocks.py:277-314)
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=3, padding=1)


class _PosePatchEmbed(nn.Module):           # embeddings.py:390-428 (RandomDropoutPatchEmbed; eval: row mask only)
    def __init__(self, in_ch: int, dim: int, p: int):
        super().__init__()
        self.patch_embedder = _PatchEmbed(in_ch, dim, p)


class PoseCondition:
    """Camera-pose conditioning of a window, handed to the backbone instead of the reference's dense
    (R, T, 180, H, W) ray-encoding tensor.  `cams` [n_cond, T, 16] f32 = per frame (fx, fy, px, py) in pixels, R^-1
    row-major, ray origin (see dfot_video_pose.camera_table); `row_map[r]` = conditioning row of backbone row r
    (history-guidance branches of one sample share a row)."""

    def __init__(self, cams: torch.Tensor, row_map: List[int], token=None):
        self.cams = cams
        self.row_map = list(row_map)
        self.token = token if token is not None else object()   # identity of the cached content

    @property
    def shape(self):
        return (len(self.row_map),) + tuple(self.cams.shape[1:])

    def index_select(self, dim: int, rows: torch.Tensor) -> "PoseCondition":
        assert dim == 0
        return PoseCondition(self.cams, [self.row_map[int(r)] for r in rows.tolist()], self.token)


class UViT3DPose(nn.Module):
    def __init__(self, cfg, x_shape, max_tokens: int, external_cond_dim: int = 0, use_causal_mask: bool = True,
                 external_cond_type: Optional[str] = None, external_cond_num_classes: Optional[int] = None, **kwargs):
        super().__init__()
        cfg = to_config(cfg)
        self.cfg = cfg
        self.x_shape = list(x_shape)
        self.max_tokens = self.temporal_length = max_tokens
        self.external_cond_type = external_cond_type
        self.external_cond_num_classes = external_cond_num_classes
        self.external_cond_dim = cfg.conditioning.dim          # overwritten by DFoTVideoPose._update_backbone_cfg
        self.use_causal_mask = use_causal_mask
        self.channels = list(cfg.channels)
        self.emb_dim = E = cfg.emb_channels
        self.patch_size = p = cfg.patch_size
        self.block_types = list(cfg.block_types)
        if any(b not in ("ResBlock", "TransformerBlock") for b in self.block_types) or cfg.pos_emb_type != "rope":
            raise NotImplementedError("dfot_b200 UViT3DPose supports ResBlock / TransformerBlock levels with "
                                      "pos_emb_type=rope (u_vit3d_pose.yaml); axial / learned_1d are unused ablations")
        if any(d != 0.0 for d, b in zip(cfg.block_dropouts, self.block_types) if b == "ResBlock"):
            raise AssertionError("Dropout is not supported in ResBlock.")
        self.num_updown_blocks = list(cfg.num_updown_blocks)
        self.num_mid_blocks = cfg.num_mid_blocks
        self.num_heads = cfg.num_heads
        self.num_levels = L = len(self.channels)
        self.is_transformers = [b != "ResBlock" for b in self.block_types]
        C, H, W = self.x_shape
        if H != W or H % (p << (L - 1)):
            raise ValueError(f"x_shape {self.x_shape} incompatible with patch {p} and {L} levels")
        self.res = [H // p // (2 ** i) for i in range(L)]
        for i, ch in enumerate(self.channels):
            if self.is_transformers[i] and (ch % self.num_heads or ch // self.num_heads not in (64, 128)):
                raise NotImplementedError(f"head_dim {ch // self.num_heads} unsupported by the attention kernel (64, 128)")
            if not self.is_transformers[i] and ch % 32:
                raise ValueError("GroupNorm(32) needs channels % 32 == 0")
        self.conditioning_dropout = cfg.external_cond_dropout

        self.noise_level_pos_embedding = _NoiseLevelEmbedding(256, E, bool(cfg.get("use_fourier_noise_embedding", False)))
        self.external_cond_embedding = _PosePatchEmbed(self.external_cond_dim, E, p)
        self.embed_input = _EmbedInput(C, self.channels[0], p)
        self.project_output = _ProjectOutput(self.channels[0], C, p)
        self.pos_embs = nn.ModuleDict({})     # RoPE tables hold no parameters/persistent buffers in the reference

        def block(level: int) -> nn.Module:
            ch = self.channels[level]
            return _TransformerBlock(ch, self.num_heads, E) if self.is_transformers[level] else _ResBlock(ch, E)

        # registration order = the reference's (u_vit3d.py:113-152: down_blocks, up_blocks, then mid_blocks): the EMA list
        # of a Lightning checkpoint is zipped with named_parameters(), so the ORDER is part of the drop-in contract
        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()
        for i in range(L - 1):
            self.down_blocks.append(nn.ModuleList([block(i) for _ in range(self.num_updown_blocks[i])]
                                                  + [_Resample(self.channels[i], self.channels[i + 1])]))
        self.mid_blocks = nn.ModuleList([block(L - 1) for _ in range(self.num_mid_blocks)])
        for u in range(L - 1):
            i = L - 2 - u
            self.up_blocks.append(nn.ModuleList([_Resample(self.channels[i + 1], self.channels[i])]
                                                + [block(i) for _ in range(self.num_updown_blocks[i])]))
        self._packed = None
        self._packed_key = None
        self._ws = {}
        self._pose_state = {}
        self.use_cuda_graph = True
        self._graphs = {}
        self._img_base = {}

    @property
    def n_tokens_per_frame(self) -> int:
        return self.x_shape[1] * self.x_shape[2] // (self.patch_size ** 2)

    # ------------------------------------------------------------------ block enumeration (execution order)
    def _blocks_in_order(self):
        """[(module, level)] of every Res/Transformer block in forward order."""
        L = self.num_levels
        out = []
        for i in range(L - 1):
            out += [(b, i) for b in list(self.down_blocks[i])[:-1]]
        out += [(b, L - 1) for b in self.mid_blocks]
        for u in range(L - 1):
            out += [(b, L - 2 - u) for b in list(self.up_blocks[u])[1:]]
        return out

    # ------------------------------------------------------------------ weight packing
    def _version_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + \
               tuple((b.data_ptr(), b._version) for b in self.buffers())

    def packed(self):
        key = self._version_key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        dev = self.embed_input.proj.weight.device
        ops.require_cuda(dev, "UViT3DPose")
        bf = lambda w: ops.cast_bf16(w.detach().float().contiguous())
        f32 = lambda b: b.detach().float().contiguous()
        conv_w = lambda c: bf(c.weight.detach().float().permute(0, 2, 3, 1))      # [Cout, 3, 3, Cin]
        C, p, E = self.x_shape[0], self.patch_size, self.emb_dim
        P = {}
        te = self.noise_level_pos_embedding.embedding
        P["t1_w"], P["t1_b"], P["t2_w"], P["t2_b"] = bf(te.linear_1.weight), f32(te.linear_1.bias), \
            bf(te.linear_2.weight), f32(te.linear_2.bias)
        if hasattr(self.noise_level_pos_embedding, "timesteps"):
            P["four_f"] = f32(self.noise_level_pos_embedding.timesteps.freqs)
            P["four_p"] = f32(self.noise_level_pos_embedding.timesteps.phases)
        # input patch embed: K = C*p*p padded to 8 (columns (c, py, px) = patchify order)
        kp = _pad8(C * p * p)
        wi = torch.zeros((self.channels[0], kp), device=dev)
        wi[:, : C * p * p] = self.embed_input.proj.weight.detach().float().reshape(self.channels[0], -1)
        P["in_w"], P["in_b"] = bf(wi), f32(self.embed_input.proj.bias)
        # output projection: ConvTranspose2d(k = s = p) as a GEMM with rows (py, px, c_out) = unpatchify order
        no = _pad8(p * p * C)
        wo = torch.zeros((no, self.channels[0]), device=dev)
        wo[: p * p * C] = self.project_output.proj.weight.detach().float().permute(2, 3, 1, 0).reshape(p * p * C, -1)
        bo = torch.zeros((no,), device=dev)
        bo[: p * p * C] = self.project_output.proj.bias.detach().float().repeat(p * p)
        P["out_w"], P["out_b"] = bf(wo), bo
        # pose patch embed, two column orders: (py, px, c) for the fused ray-patch kernel, (c, py, px) for dense input
        wp = self.external_cond_embedding.patch_embedder.proj.weight.detach().float()     # [E, Cc, p, p]
        P["pose_w_fast"] = bf(wp.permute(0, 2, 3, 1).reshape(E, -1))
        P["pose_w_dense"] = bf(wp.reshape(E, -1))
        P["pose_b"] = f32(self.external_cond_embedding.patch_embedder.proj.bias)
        # FiLM emb layers: per-frame part of all blocks in ONE GEMM (bias included here), per-pixel part per block
        mods_w, mods_b, blocks = [], [], []
        col = 0
        for blk, lvl in self._blocks_in_order():
            ch = self.channels[lvl]
            d = dict(level=lvl, kind=blk.kind, col=col)
            if blk.kind == "res":
                ew = blk.emb_layer.weight.detach().float().reshape(2 * ch, E)
                eb = blk.emb_layer.bias
                d.update(gn1_w=f32(blk.in_layers[0].weight), gn1_b=f32(blk.in_layers[0].bias),
                         conv1_w=conv_w(blk.in_layers[2]), conv1_b=f32(blk.in_layers[2].bias),
                         gn2_w=f32(blk.out_norm.weight), gn2_b=f32(blk.out_norm.bias),
                         conv2_w=conv_w(blk.out_rest[1]), conv2_b=f32(blk.out_rest[1].bias))
            else:
                ew = blk.norm.emb_layer.weight.detach().float()
                eb = blk.norm.emb_layer.bias
                fw, fb = blk.fused_attn_mlp_proj.weight.detach().float(), blk.fused_attn_mlp_proj.bias.detach().float()
                # attn_out and mlp_out both add into the residual stream: ONE GEMM over the concatenated operand
                # [attention | SiLU(mlp_h)] (K = 5C) with concatenated weights and summed biases
                out_w = torch.cat([blk.attn_out.weight.detach().float(), blk.mlp_out[2].weight.detach().float()], 1)
                d.update(norm_w=f32(blk.norm.norm.weight), qkv_w=bf(fw[: 3 * ch]), qkv_b=fb[: 3 * ch].contiguous(),
                         mlp_w=bf(fw[3 * ch:]), mlp_b=fb[3 * ch:].contiguous(), qn_w=f32(blk.q_norm.weight),
                         kn_w=f32(blk.k_norm.weight), out_w=bf(out_w),
                         # QK-normalised attention has bounded logits: |q.k| * scale * log2e <= sqrt(d) max|w_q| max|w_k| log2e
                         # (Cauchy-Schwarz on RMS-normalised, RoPE-rotated vectors; 5 % slack for bf16 rounding)
                         score_bound=1.05 * LOG2E * math.sqrt(ch // self.num_heads) *
                         float(blk.q_norm.weight.detach().abs().max()) * float(blk.k_norm.weight.detach().abs().max()),
                         out_b=(blk.attn_out.bias.detach().float() + blk.mlp_out[2].bias.detach().float()).contiguous())
            d["emb_w"] = bf(ew)
            mods_w.append(ew)
            mods_b.append(eb.detach().float())
            col += 2 * ch
            blocks.append(d)
        P["blocks"] = blocks
        P["mod_w"] = bf(torch.cat(mods_w, 0))
        P["mod_b"] = torch.cat(mods_b, 0).contiguous()
        P["down"] = [dict(w=conv_w(self.down_blocks[i][-1].conv), b=f32(self.down_blocks[i][-1].conv.bias))
                     for i in range(self.num_levels - 1)]
        P["up"] = [dict(w=conv_w(self.up_blocks[u][0].conv), b=f32(self.up_blocks[u][0].conv.bias))
                   for u in range(self.num_levels - 1)]
        P["rope"] = {i: rope_cos_sin_table(self.channels[i] // self.num_heads,
                                           (self.temporal_length, self.res[i], self.res[i])).to(dev)
                     for i in range(self.num_levels) if self.is_transformers[i]}
        self._packed, self._packed_key = P, key
        self._pose_state = {}          # cached pose modulation depends on the weights
        return P

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, R: int, dev, out_dtype):
        key = (R, str(dev), out_dtype)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        T, E, p, C = self.temporal_length, self.emb_dim, self.patch_size, self.x_shape[0]
        n = R * T
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        bf, f32 = torch.bfloat16, torch.float32
        n_mod = sum(2 * self.channels[l] for _, l in self._blocks_in_order())
        M0 = n * self.res[0] ** 2
        ws = dict(feat=e((n, 256), bf), e1=e((n, E), bf), emb=e((n, E), bf), mod=e((n, n_mod), f32),
                  patches=torch.zeros((M0, _pad8(C * p * p)), dtype=bf, device=dev), sums=e((n, 32, 3), torch.float64), sums_h=e((n, 32, 3), torch.float64),
                  x0_16=e((M0, self.channels[0]), bf), tok=e((M0, _pad8(p * p * C)), f32),
                  out=e((R, T, *self.x_shape), out_dtype), img_map=e((n,), torch.int32), lv=[])
        for i, ch in enumerate(self.channels):
            M = n * self.res[i] ** 2
            d = dict(x=e((M, ch), f32), a16=e((M, ch), bf))
            if self.is_transformers[i]:
                d.update(qkv=e((M, 3 * ch), bf), cat=e((M, 5 * ch), bf))     # cat = [attention out | SiLU(mlp_h)]
            else:
                d.update(h16=e((M, ch), bf))
            if i + 1 < self.num_levels:
                cn = self.channels[i + 1]
                d.update(pool=e((M // 4, ch), bf), after=e((M // 4, cn), f32), diff=e((M // 4, cn), bf),
                         low=e((M // 4, ch), f32))
            ws["lv"].append(d)
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ per-window camera-pose modulation cache
    def _pose_buffers(self, n_cond: int, dev):
        key = (n_cond, str(dev))
        st = self._pose_state.get(key)
        if st is None:
            T, E = self.temporal_length, self.emb_dim
            e = lambda shape: torch.empty(shape, dtype=torch.bfloat16, device=dev)
            st = dict(token=None, n_cond=n_cond,
                      pe=[e((n_cond * T * self.res[i] ** 2, E)) for i in range(self.num_levels)],
                      cache=[e((n_cond * T * self.res[l] ** 2, 2 * self.channels[l])) for _, l in self._blocks_in_order()])
            self._pose_state[key] = st
        return st

    def _fill_pose_cache(self, st, patches: torch.Tensor, weight: torch.Tensor):
        """patches [n_cond*T*res0^2, K] bf16 PatchEmbed rows → every block's per-pixel [scale | shift] (bf16)."""
        Pk = self.packed()
        T, E, n_img = self.temporal_length, self.emb_dim, st["n_cond"] * self.temporal_length
        ops.gemm_bf16(patches, weight, st["pe"][0], ops.EPI_BF16, bias=Pk["pose_b"])
        for i in range(1, self.num_levels):
            ops.avgpool2x2(st["pe"][i - 1], st["pe"][i], n_img, self.res[i - 1], self.res[i - 1], E)
        for buf, bw in zip(st["cache"], Pk["blocks"]):
            ops.gemm_bf16(st["pe"][bw["level"]], bw["emb_w"], buf, ops.EPI_BF16)

    def prepare_pose(self, cond: PoseCondition):
        """Make the HBM-resident modulation cache hold `cond` (no-op when it already does)."""
        dev = cond.cams.device
        ops.require_cuda(dev, "UViT3DPose.prepare_pose")
        self.packed()
        n_cond, T = cond.cams.shape[:2]
        if T != self.temporal_length:
            raise ValueError(f"pose condition has {T} frames, the backbone expects {self.temporal_length}")
        st = self._pose_buffers(n_cond, dev)
        if st["token"] is cond.token:
            return st
        from dfot_b200.algorithms.dfot.dfot_video_pose import ray_freq_scale
        p, res = self.patch_size, self.x_shape[1]
        n_freq = self.external_cond_dim // 12
        g = res // p
        patches = torch.empty((n_cond * T * g * g, p * p * self.external_cond_dim), dtype=torch.bfloat16, device=dev)
        ops.pose_ray_patches(cond.cams.reshape(n_cond * T, 16).float().contiguous(), ray_freq_scale(n_freq).to(dev),
                             patches, n_cond * T, res, p)
        self._fill_pose_cache(st, patches, self._packed["pose_w_fast"])
        st["token"] = cond.token
        return st

    def _prepare_dense(self, external_cond: torch.Tensor):
        """Reference-style dense (R, T, Cc, H, W) conditioning: every row is its own conditioning row."""
        R, T = external_cond.shape[:2]
        Cc, p, res = self.external_cond_dim, self.patch_size, self.x_shape[1]
        st = self._pose_buffers(R, external_cond.device)
        g = res // p
        patches = torch.empty((R * T * g * g, Cc * p * p), dtype=torch.bfloat16, device=external_cond.device)
        ops.patchify_bf16(external_cond.float().contiguous(), patches, R * T, Cc, res, res, p)
        self._fill_pose_cache(st, patches, self.packed()["pose_w_dense"])
        st["token"] = None
        return st

    # ------------------------------------------------------------------ forward
    def input_buffer(self, R: int, T: int, dtype, device) -> torch.Tensor:
        key = ("in", R, T, dtype, str(device))
        buf = self._ws.get(key)
        if buf is None:
            buf = torch.empty((R, T, *self.x_shape), dtype=dtype, device=device)
            self._ws[key] = buf
        return buf

    @torch.no_grad()
    def forward(self, x: torch.Tensor, noise_levels: torch.Tensor, external_cond=None,
                external_cond_mask: Optional[torch.Tensor] = None, out_dtype=torch.float32) -> torch.Tensor:
        """x [R,T,C,H,W] f32|bf16; noise_levels [R,T] int64 | f32; external_cond: PoseCondition (fast path) or the
        reference's dense (R,T,180,H,W) ray-encoding tensor; external_cond_mask [R] bool (True = drop the pose)."""
        ops.require_cuda(x.device, "UViT3DPose.forward")
        assert x.shape[1] == self.temporal_length, \
            f"Temporal length of U-ViT is set to {self.temporal_length}, but input has temporal length {x.shape[1]}."
        assert external_cond is not None, "External condition (camera pose) is required for U-ViT3DPose model."
        R, T = x.shape[:2]
        dev = x.device
        if isinstance(external_cond, PoseCondition):
            st = self.prepare_pose(external_cond)
            row_map = external_cond.row_map
        else:
            st = self._prepare_dense(external_cond)
            row_map = list(range(R))
        if len(row_map) != R:
            raise ValueError(f"conditioning covers {len(row_map)} rows, x has {R}")
        # image -> row of the pose cache (constant over the steps of a window: built and uploaded once per row map)
        map_key = (tuple(row_map), T, str(dev))
        if self._img_base.get("key") != map_key:
            rm = torch.tensor(row_map, dtype=torch.int32)
            self._img_base = {"key": map_key,
                              "base": (rm[:, None] * T + torch.arange(T, dtype=torch.int32)[None, :]).to(dev)}
        base = self._img_base["base"]
        levels = noise_levels if noise_levels.dtype in (torch.int64, torch.float32) else noise_levels.float()
        graph_ok = self.use_cuda_graph and x.is_cuda and isinstance(external_cond, PoseCondition) and \
            not torch.cuda.is_current_stream_capturing()
        if not graph_ok:
            ws = self._workspace(R, dev, out_dtype)
            self._set_img_map(ws["img_map"], base, external_cond_mask)
            return self._forward_impl(x, levels, st, ws)
        sig = (R, x.dtype, levels.dtype, out_dtype, str(dev), st["n_cond"])
        g = self._graphs.get(sig)
        if g is None:        # first call with this signature runs eagerly (warms lazy kernel attributes)
            self._graphs[sig] = {"graph": None, "key": self._version_key()}
            ws = self._workspace(R, dev, out_dtype)
            self._set_img_map(ws["img_map"], base, external_cond_mask)
            return self._forward_impl(x, levels, st, ws)
        if g["key"] != self._version_key():
            g["graph"], g["key"] = None, self._version_key()
        ws = self._workspace(R, dev, out_dtype)
        xin = self.input_buffer(R, T, x.dtype, dev)
        if x.data_ptr() != xin.data_ptr():
            xin.copy_(x)
        if g["graph"] is None:
            g["levels"] = torch.empty_like(levels)
        g["levels"].copy_(levels)
        self._set_img_map(ws["img_map"], base, external_cond_mask)
        if g["graph"] is None:
            graph = torch.cuda.CUDAGraph()
            n0 = _abi.launch_count()
            with torch.cuda.graph(graph):
                g["out"] = self._forward_impl(xin, g["levels"], st, ws)
            g["graph"], g["kernels"] = graph, _abi.launch_count() - n0
        g["graph"].replay()
        ops.count_replayed_launches(g["kernels"])
        return g["out"]

    @staticmethod
    def _set_img_map(img_map: torch.Tensor, base: torch.Tensor, mask: Optional[torch.Tensor]):
        if mask is not None:
            base = torch.where(mask.to(base.device).bool()[:, None], torch.full_like(base, -1), base)
        img_map.copy_(base.reshape(-1))

    @torch.no_grad()
    def _forward_impl(self, x: torch.Tensor, levels: torch.Tensor, st, ws) -> torch.Tensor:
        R, T = x.shape[:2]
        C, H, W = self.x_shape
        p, L, n = self.patch_size, self.num_levels, R * T
        Pk = self.packed()
        lv, mod, img_map, sums, sums_h = ws["lv"], ws["mod"], ws["img_map"], ws["sums"], ws["sums_h"]
        x = x.contiguous()
        levels = levels.contiguous()

        # --- input patches → level-0 feature map (channel-last)
        ops.patchify_bf16(x, ws["patches"], n, C, H, W, p)
        ops.gemm_bf16(ws["patches"], Pk["in_w"], lv[0]["x"], ops.EPI_F32, bias=Pk["in_b"])
        # --- per-frame noise-level embedding and its FiLM contribution for ALL blocks (one GEMM)
        ops.noise_features(levels, ws["feat"], Pk.get("four_f"), Pk.get("four_p"))
        ops.gemm_bf16(ws["feat"], Pk["t1_w"], ws["e1"], ops.EPI_SILU_BF16, bias=Pk["t1_b"])
        ops.gemm_bf16(ws["e1"], Pk["t2_w"], ws["emb"], ops.EPI_BF16, bias=Pk["t2_b"])
        ops.gemm_bf16(ws["emb"], Pk["mod_w"], mod, ops.EPI_F32, bias=Pk["mod_b"])

        blocks = iter(zip(Pk["blocks"], st["cache"]))

        def run_level(i: int, count: int, src: torch.Tensor) -> torch.Tensor:
            """Blocks of level i; the first one reads `src` and writes the level buffer, the rest run in place."""
            w, ch = lv[i], self.channels[i]
            g = self.res[i]
            HW = g * g
            dst = w["x"]
            # GroupNorm statistics ride on the conv epilogues (of the values as stored); only a level's first ResBlock
            # needs a stand-alone statistics pass over its input.  (HW % 32 == 0 is what the epilogue path needs.)
            fused_stats = HW % 32 == 0 and ch // 32 in (1, 2, 4, 8, 16, 32)
            have_stats = False
            for b in range(count):
                bw, cache = next(blocks)
                assert bw["level"] == i
                sc, sh = bw["col"], bw["col"] + ch
                if bw["kind"] == "res":
                    if not have_stats:
                        ops.groupnorm_stats(src, sums, n, HW, ch)
                    ops.groupnorm_silu_bf16(src, sums, bw["gn1_w"], bw["gn1_b"], w["a16"], n, HW, ch)
                    ops.conv3x3_bf16(w["a16"].view(n, g, g, ch), bw["conv1_w"], w["h16"], ops.EPI_BF16, bias=bw["conv1_b"],
                                     gn_sums=sums_h if fused_stats else None)
                    if not fused_stats:
                        ops.groupnorm_stats(w["h16"], sums_h, n, HW, ch)
                    ops.groupnorm_silu_bf16(w["h16"], sums_h, bw["gn2_w"], bw["gn2_b"], w["a16"], n, HW, ch, mod_img=mod,
                                            scale_col=sc, shift_col=sh, mod_pix=cache, img_map=img_map)
                    have_stats = fused_stats and b + 1 < count      # statistics of the block's output for the next block
                    ops.conv3x3_bf16(w["a16"].view(n, g, g, ch), bw["conv2_w"], dst, ops.EPI_RESID_F32,
                                     bias=bw["conv2_b"], resid=src, gn_sums=sums if have_stats else None)
                else:
                    dh = ch // self.num_heads
                    Ntok = T * HW
                    ops.rmsnorm_film_bf16(src, bw["norm_w"], mod, sc, sh, HW, w["a16"], mod_pix=cache, img_map=img_map)
                    if ch >= _FUSE_QKNORM_MIN_CH:
                        # wide levels (K >= 1024: the GEMM is MMA-bound, its epilogue has slack): q/k RMSNorm(head_dim)
                        # + RoPE-3D + softmax scale ride on the QKV GEMM's epilogue, from the fp32 accumulators
                        ops.gemm_bf16(w["a16"], bw["qkv_w"], w["qkv"], ops.EPI_QKNORM_ROPE_BF16, bias=bw["qkv_b"],
                                      rope_cs=Pk["rope"][i], tokens_per_sample=Ntok, model_dim=ch, head_dim=dh,
                                      q_scale=LOG2E / math.sqrt(dh), qn_w=bw["qn_w"], kn_w=bw["kn_w"])
                    else:
                        # narrow levels (K = 576 at RE10K): the GEMM is epilogue-bound, a separate HBM-bound pass is
                        # cheaper than a heavier epilogue (measured: 134.2 vs 132.7 NFE/s in r01; 11.01 vs 10.93 frames/s
                        # after the r02 issue fix)
                        ops.gemm_bf16(w["a16"], bw["qkv_w"], w["qkv"], ops.EPI_BF16, bias=bw["qkv_b"])
                        ops.qk_norm_rope(w["qkv"], bw["qn_w"], bw["kn_w"], Pk["rope"][i], Ntok, self.num_heads, dh,
                                         LOG2E / math.sqrt(dh))
                    ops.gemm_bf16(w["a16"], bw["mlp_w"], w["cat"][:, ch:], ops.EPI_SILU_BF16, bias=bw["mlp_b"])
                    ops.attention(w["qkv"], w["cat"][:, :ch], R, Ntok, self.num_heads, dh, score_bound=bw["score_bound"])
                    ops.gemm_bf16(w["cat"], bw["out_w"], dst, ops.EPI_RESID_F32, bias=bw["out_b"], resid=src)
                src = dst
            return dst

        # --- down path
        cur = lv[0]["x"]
        for i in range(L - 1):
            cur = run_level(i, self.num_updown_blocks[i], cur)          # == lv[i]["x"] (hs_before)
            g = self.res[i]
            ops.avgpool2x2(cur, lv[i]["pool"], n, g, g, self.channels[i])
            ops.conv3x3_bf16(lv[i]["pool"].view(n, g // 2, g // 2, self.channels[i]), Pk["down"][i]["w"],
                             lv[i]["after"], ops.EPI_F32, bias=Pk["down"][i]["b"])
            cur = lv[i]["after"]                                        # hs_after: kept intact, next level writes lv[i+1].x
            if self.num_updown_blocks[i + 1] == 0 if i + 1 < L - 1 else self.num_mid_blocks == 0:
                lv[i + 1]["x"].copy_(cur)
                cur = lv[i + 1]["x"]
        # --- middle
        cur = run_level(L - 1, self.num_mid_blocks, cur)
        # --- up path
        for u in range(L - 1):
            i = L - 2 - u
            g = self.res[i]
            ops.sub_bf16(cur, lv[i]["after"], lv[i]["diff"])
            ops.conv3x3_bf16(lv[i]["diff"].view(n, g // 2, g // 2, self.channels[i + 1]), Pk["up"][u]["w"], lv[i]["low"],
                             ops.EPI_F32, bias=Pk["up"][u]["b"])
            ops.upsample2x_add(lv[i]["low"], lv[i]["x"], lv[i]["x"], n, g, g, self.channels[i])
            cur = run_level(i, self.num_updown_blocks[i], lv[i]["x"])
        # --- output projection + unpatchify
        ops.cast_bf16(cur, ws["x0_16"])
        ops.gemm_bf16(ws["x0_16"], Pk["out_w"], ws["tok"], ops.EPI_F32, bias=Pk["out_b"])
        ops.unpatchify(ws["tok"], ws["out"], n, C, H, W, p)
        return ws["out"]
