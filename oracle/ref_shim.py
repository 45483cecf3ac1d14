"""TEST INFRASTRUCTURE — import shim for the *real* reference (/root/reference).

Used by ``oracle/make_goldens*.py`` (authoring container, /root/reference) to (a) generate
the committed golden vectors under ``tests/golden/`` and (b) pin the CPU restatement in
``oracle/`` against the executed reference; and by ``bench.py --impl reference`` on the
GPU box, where the reference's files come from the git-ignored copy ``oracle/_ref/``
made by ``oracle/build_ref.py``.  Nothing in the product package imports this.

The reference depends on packages that are absent here (omegaconf, timm,
diffusers, lightning, accelerate, rotary_embedding_torch, roma, matplotlib,
colorama, wandb ...).  We pre-seed ``sys.modules`` with minimal stand-ins whose
semantics restate the pinned versions in the reference's ``requirements.txt``
(timm==1.0.17 PatchEmbed/Mlp, diffusers==0.32.2 TimestepEmbedding/LabelEmbedding,
rotary_embedding_torch==0.8.6 rotate_half) at the reference's own call sites:
  algorithms/dfot/backbones/dit/dit3d.py:49-55      (PatchEmbed)
  algorithms/dfot/backbones/dit/dit_blocks.py:469-473 (Mlp)
  algorithms/dfot/backbones/modules/embeddings.py:8,84,215 (rotate_half, TimestepEmbedding)
  algorithms/dfot/backbones/base_backbone.py:47-51  (LabelEmbedding)
"""
import os
import sys
import types

import torch
from torch import nn

_SHIPPED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")   # oracle/build_ref.py


def _find_reference() -> str:
    """$DFOT_REFERENCE_ROOT, else /root/reference (authoring container), else the unmodified copy of the sampling
    path's files that oracle/build_ref.py ships to the GPU box (git-ignored oracle/_ref/)."""
    env = os.environ.get("DFOT_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _SHIPPED):
        if os.path.isdir(os.path.join(cand, "algorithms", "dfot")):
            return cand
    return "/root/reference"


REF = _find_reference()


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "algorithms", "dfot"))


def _mod(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


class DictConfig(dict):
    """attr-dict stand-in for omegaconf.DictConfig"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def to_dc(x):
    if isinstance(x, dict):
        return DictConfig({k: to_dc(v) for k, v in x.items()})
    if isinstance(x, (list, tuple)):
        return [to_dc(v) for v in x]
    return x


def _plain(x):
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    return x


_installed = False


def install():
    """Install all stand-ins and make the reference importable. Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REF}")
    _installed = True

    # ---- omegaconf
    oc = _mod("omegaconf")

    class OmegaConf:
        @staticmethod
        def to_container(c, resolve=True):
            return _plain(c)

        @staticmethod
        def create(x):
            return to_dc(x)

    oc.DictConfig = DictConfig
    oc.OmegaConf = OmegaConf
    oc.open_dict = None

    # ---- timm (PatchEmbed = strided conv + flatten; Mlp = fc1-act-fc2)
    _mod("timm")
    _mod("timm.models")
    tv = _mod("timm.models.vision_transformer")

    class PatchEmbed(nn.Module):
        def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768,
                     norm_layer=None, flatten=True, bias=True, **kw):
            super().__init__()
            if isinstance(img_size, int):
                img_size = (img_size, img_size)
            self.patch_size = (patch_size, patch_size)
            self.img_size = tuple(img_size)
            self.grid_size = tuple(s // patch_size for s in img_size)
            self.num_patches = self.grid_size[0] * self.grid_size[1]
            self.flatten = flatten
            self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size,
                                  stride=patch_size, bias=bias)
            self.norm = nn.Identity()

        def forward(self, x):
            x = self.proj(x)
            if self.flatten:
                x = x.flatten(2).transpose(1, 2)
            return self.norm(x)

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None,
                     act_layer=nn.GELU, norm_layer=None, bias=True, drop=0.0, use_conv=False):
            super().__init__()
            out_features = out_features or in_features
            hidden_features = hidden_features or in_features
            self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
            self.act = act_layer()
            self.drop1 = nn.Dropout(drop)
            self.norm = nn.Identity()
            self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
            self.drop2 = nn.Dropout(drop)

        def forward(self, x):
            return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))

    tv.PatchEmbed = PatchEmbed
    tv.Mlp = Mlp

    # ---- diffusers
    _mod("diffusers")
    _mod("diffusers.models")
    de = _mod("diffusers.models.embeddings")

    class TimestepEmbedding(nn.Module):
        def __init__(self, in_channels, time_embed_dim, act_fn="silu", out_dim=None):
            super().__init__()
            self.linear_1 = nn.Linear(in_channels, time_embed_dim)
            self.act = nn.SiLU()
            self.linear_2 = nn.Linear(time_embed_dim, out_dim or time_embed_dim)

        def forward(self, sample, condition=None):
            return self.linear_2(self.act(self.linear_1(sample)))

    class LabelEmbedding(nn.Module):
        def __init__(self, num_classes, hidden_size, dropout_prob):
            super().__init__()
            self.embedding_table = nn.Embedding(num_classes + int(dropout_prob > 0), hidden_size)
            self.num_classes = num_classes
            self.dropout_prob = dropout_prob

        def forward(self, labels, force_drop_ids=None):
            return self.embedding_table(labels)

    de.TimestepEmbedding = TimestepEmbedding
    de.LabelEmbedding = LabelEmbedding

    # ---- rotary_embedding_torch
    _mod("rotary_embedding_torch")
    rr = _mod("rotary_embedding_torch.rotary_embedding_torch")

    def rotate_half(x):
        x = x.reshape(*x.shape[:-1], -1, 2)
        x1, x2 = x.unbind(dim=-1)
        return torch.stack((-x2, x1), dim=-1).flatten(-2)

    rr.rotate_half = rotate_half

    # ---- lightning / accelerate / plotting / misc
    L = _mod("lightning")
    LP = _mod("lightning.pytorch")
    L.pytorch = LP

    class LightningModule(nn.Module):
        def __init__(self):
            super().__init__()
            self._trainer = None

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        @property
        def trainer(self):
            return self._trainer

        def log(self, *a, **k):
            raise AttributeError

        def log_dict(self, *a, **k):
            pass

    LP.LightningModule = LightningModule
    u = _mod("lightning.pytorch.utilities")
    u.grad_norm = lambda *a, **k: {}
    ut = _mod("lightning.pytorch.utilities.types")
    ut.STEP_OUTPUT = object
    rz = _mod("lightning.pytorch.utilities.rank_zero")

    def rank_zero_only(fn):
        return fn

    rank_zero_only.rank = 0
    rz.rank_zero_only = rank_zero_only
    _mod("lightning.pytorch.loggers")
    lgl = _mod("lightning.pytorch.loggers.logger")
    lgl.Logger = object
    _mod("lightning_utilities")
    _mod("lightning_utilities.core")
    lua = _mod("lightning_utilities.core.apply_func")
    lua.apply_to_collection = lambda data, dtype, fn: fn(data)
    acc = _mod("accelerate")
    acc.Accelerator = object
    tr = _mod("transformers")
    tr.get_scheduler = lambda *a, **k: None
    mpl = _mod("matplotlib")
    plt = _mod("matplotlib.pyplot")
    mpl.pyplot = plt
    plt.Figure = object
    plt.Axes = object
    col = _mod("colorama")

    class _F:
        CYAN = ""
        RESET = ""

    col.Fore = _F
    from . import roma_restatement           # roma 1.5.2.1 is absent: its three functions on the pose path, restated
    roma_restatement.install(_mod("roma"))   # (pinned against scipy, tests/test_pose_quaternions.py)
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            _mod("wandb")

    # ---- reference packages: path-only (skip their heavy __init__)
    for pkg in ["algorithms", "algorithms.dfot", "algorithms.common", "algorithms.dfot.backbones", "utils"]:
        m = _mod(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split("."))]
    _mod("algorithms.common.metrics")
    mv = _mod("algorithms.common.metrics.video")
    mv.VideoMetric = object
    mv.SharedVideoMetricModelRegistry = object
    ah = _mod("algorithms.common.attn_hook")
    for n in ["register_hooks", "clear_hooks", "save_attention_maps"]:
        setattr(ah, n, lambda *a, **k: None)
    ah.attn_maps = {}
    v = _mod("algorithms.vae")
    # (latent configurations with `latent.type: online` construct their VAE in __init__; the sampling path on latents never
    #  calls it, and no VAE checkpoint exists offline: an empty module stands in)
    for n in ["ImageVAE", "VideoVAE", "MyAutoencoderDC", "AutoencoderKL", "TiTok_KL"]:
        setattr(v, n, type(n, (), {"from_pretrained": classmethod(lambda cls, *a, **k: nn.Identity())}))
    ulog = _mod("utils.logging_utils")
    ulog.log_video = lambda *a, **k: None

    if REF not in sys.path:
        sys.path.insert(0, REF)
    b = sys.modules["algorithms.dfot.backbones"]
    from algorithms.dfot.backbones.dit.dit3d import DiT3D
    from algorithms.dfot.backbones.u_vit.u_vit3d_pose import UViT3DPose
    for n, c in dict(Unet3D=None, DiT3D=DiT3D, DiT3DPose=None, UViT3D=None, UViT3DPose=UViT3DPose,
                     FARDiT=None, DIT1D=None, DifferenceDiT3D=None).items():
        setattr(b, n, c)


def rerandomize_zero_params(module: nn.Module, seed: int, std: float = 0.02) -> None:
    """A fresh reference model outputs exactly 0 (all output layers are
    zero-initialised).  Re-draw every all-zero parameter with
    N(0, std) from a seeded generator, in named_parameters() order."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, p in module.named_parameters():
            if p.numel() > 0 and bool((p == 0).all()):
                p.copy_(torch.randn(p.shape, generator=g) * std)
