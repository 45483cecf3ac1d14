"""ctypes binding of libdfot_b200.so (the C ABI declared in include/dfot_b200.h)."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFOT_B200_LIB") or os.path.join(HERE, "libdfot_b200.so")   # env: A/B builds of the kernels

F32, BF16, I64 = 0, 1, 2
EPI_F32, EPI_BF16, EPI_GELU_BF16, EPI_SILU_BF16, EPI_GATE_RESID_F32, EPI_QKV_ROPE_BF16, EPI_RESID_F32, \
    EPI_QKNORM_ROPE_BF16, EPI_GATE_LNRESID_F32 = range(9)

SYMBOLS = [
    "dfot_abi_version", "dfot_last_error", "dfot_launch_count", "dfot_sampler_step_hg", "dfot_adaln_layernorm",
    "dfot_adaln_layernorm_stats",
    "dfot_gemm_bf16", "dfot_attention", "dfot_attention_strided", "dfot_attention_bounded", "dfot_noise_features", "dfot_silu_sum_bf16", "dfot_patchify_bf16",
    "dfot_unpatchify", "dfot_cast_bf16", "dfot_patch_mix_bf16", "dfot_patch_expand_gate_resid", "dfot_gemm_bf16_splitk", "dfot_splitk_gate_resid_adaln", "dfot_set_latency_mode", "dfot_get_latency_mode", "dfot_conv3x3_bf16", "dfot_conv3d_causal_bf16", "dfot_groupnorm_stats", "dfot_groupnorm_silu_bf16",
    "dfot_rmsnorm_film_bf16", "dfot_qk_norm_rope", "dfot_avgpool2x2", "dfot_sub_bf16", "dfot_upsample2x_add",
    "dfot_pose_ray_patches", "dfot_groupnorm_stats_strided", "dfot_groupnorm_apply_bf16", "dfot_vae_upsample2x_bf16",
    "dfot_vae_fill_pad_frames", "dfot_softmax_rows_bf16", "dfot_upsample2x_nearest_bf16",
    "dfot_relu_bf16", "dfot_pixel_shuffle2x", "dfot_linear_attention_relu", "dfot_dwconv3x3_glu_bf16", "dfot_rmsnorm_affine",
]


class FrameUpdate(Structure):
    _fields_ = [("a", c_float), ("b", c_float), ("sigma", c_float), ("w", c_float), ("clip", c_float),
                ("generate", c_int32)]


class FramePrepare(Structure):
    _fields_ = [("mode", c_int32), ("noise_row", c_int32), ("qa", c_float), ("qb", c_float)]


class GemmEpilogue(Structure):
    _fields_ = [("bias", c_void_p), ("resid", c_void_p), ("ld_resid", c_int64), ("gate", c_void_p),
                ("ld_gate", c_int64), ("tokens_per_frame", c_int64), ("rope_cs", c_void_p),
                ("tokens_per_sample", c_int64), ("model_dim", c_int64), ("head_dim", c_int64),
                ("q_scale", c_float), ("qn_w", c_void_p), ("kn_w", c_void_p), ("qk_eps", c_float),
                ("gn_sums", c_void_p), ("gn_rows_per_img", c_int64), ("gn_groups", c_int64),
                ("gn_eps", c_float), ("ln_stats", c_void_p), ("ln_shift", c_void_p), ("ln_scale", c_void_p)]


_lib = None


def lib() -> ctypes.CDLL:
    """Load the library.  It is (re)built first whenever it is missing or was built from other sources than the ones in
    the tree (content hash of csrc/ + include/, see build.source_hash) — a stale .so would mis-marshal the ctypes structs
    silently.  Fails loudly when that is needed and impossible: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("DFOT_B200_LIB"):          # (an explicitly pinned A/B build is loaded as it is)
        from . import build as B
        if not B.is_current():
            try:
                B.build()
            except Exception as e:  # noqa: BLE001
                what = "is missing" if not os.path.exists(LIB_PATH) else "is stale (built from different sources)"
                raise RuntimeError(
                    f"dfot_b200: CUDA extension {LIB_PATH} {what} and could not be built ({e}). "
                    "Run `python __graft_entry__.py build` on a machine with nvcc; there is no CPU fallback.") from e
    L = ctypes.CDLL(LIB_PATH)
    L.dfot_abi_version.restype = c_int
    L.dfot_last_error.restype = c_char_p
    L.dfot_launch_count.restype = c_int64
    vp, i64, i = c_void_p, c_int64, c_int
    L.dfot_sampler_step_hg.argtypes = [vp, vp, i, vp, i, vp, vp, vp, vp, vp, i64, i64, i64, i64, vp]
    L.dfot_adaln_layernorm.argtypes = [vp, vp, i64, i64, i64, vp, vp, i64, i64, i64, c_float, vp]
    L.dfot_adaln_layernorm_stats.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, i64, i64, i64, c_float, vp]
    L.dfot_gemm_bf16.argtypes = [vp, i64, vp, i64, vp, i64, i64, i64, i64, i, POINTER(GemmEpilogue), vp]
    L.dfot_attention.argtypes = [vp, vp, i64, i64, i64, i64, vp]
    L.dfot_attention_strided.argtypes = [vp, vp, i64, i64, i64, i64, i64, vp]
    L.dfot_attention_bounded.argtypes = [vp, vp, i64, i64, i64, i64, i64, c_float, vp]
    L.dfot_noise_features.argtypes = [vp, i, vp, vp, vp, i64, i64, vp]
    L.dfot_silu_sum_bf16.argtypes = [vp, vp, vp, i64, vp, i64, i64, vp]
    L.dfot_patchify_bf16.argtypes = [vp, i, vp, i64, i64, i64, i64, i64, i64, vp]
    L.dfot_unpatchify.argtypes = [vp, i64, vp, i, i64, i64, i64, i64, i64, vp]
    L.dfot_cast_bf16.argtypes = [vp, vp, i64, vp]
    L.dfot_set_latency_mode.argtypes = [i]
    L.dfot_get_latency_mode.argtypes = []
    L.dfot_gemm_bf16_splitk.argtypes = [vp, i64, vp, i64, vp, i64, i64, i64, i64, vp]
    L.dfot_splitk_gate_resid_adaln.argtypes = [vp, i64, vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, i64, i64, i64, c_float, vp]
    L.dfot_patch_mix_bf16.argtypes = [vp, vp, vp, i64, i64, i64, i64, i64, vp]
    L.dfot_patch_expand_gate_resid.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, vp]
    L.dfot_conv3x3_bf16.argtypes = [vp, vp, vp, i64, i64, i64, i64, i64, i64, i, POINTER(GemmEpilogue), vp]
    L.dfot_conv3d_causal_bf16.argtypes = [vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, i, POINTER(GemmEpilogue), vp]
    L.dfot_groupnorm_stats.argtypes = [vp, i, vp, i64, i64, i64, i64, c_float, vp]
    L.dfot_groupnorm_stats_strided.argtypes = [vp, i, vp, i64, i64, i64, i64, i64, c_float, vp]
    L.dfot_groupnorm_apply_bf16.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i, vp]
    L.dfot_vae_upsample2x_bf16.argtypes = [vp, vp, i64, i64, i64, i64, i64, i, vp]
    L.dfot_vae_fill_pad_frames.argtypes = [vp, i64, i64, i64, vp]
    L.dfot_upsample2x_nearest_bf16.argtypes = [vp, vp, i64, i64, i64, i64, vp]
    L.dfot_softmax_rows_bf16.argtypes = [vp, i64, vp, i64, i64, i64, c_float, vp]
    L.dfot_groupnorm_silu_bf16.argtypes = [vp, i, vp, vp, vp, vp, i64, i64, i64, vp, vp, vp, i64, i64, i64, i64, vp]
    L.dfot_rmsnorm_film_bf16.argtypes = [vp, vp, c_float, vp, i64, i64, i64, vp, vp, vp, i64, i64, i64, vp]
    L.dfot_qk_norm_rope.argtypes = [vp, i64, vp, vp, c_float, vp, i64, i64, i64, i64, c_float, vp]
    L.dfot_avgpool2x2.argtypes = [vp, i, vp, i, i64, i64, i64, i64, vp]
    L.dfot_sub_bf16.argtypes = [vp, vp, vp, i64, vp]
    L.dfot_upsample2x_add.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp]
    L.dfot_pose_ray_patches.argtypes = [vp, vp, i64, vp, i64, i64, i64, i64, vp]
    L.dfot_relu_bf16.argtypes = [vp, i64, vp]
    L.dfot_pixel_shuffle2x.argtypes = [vp, i64, vp, i64, i64, vp, vp, i64, i64, i64, i64, vp]
    L.dfot_linear_attention_relu.argtypes = [vp, i64, vp, i64, i64, i64, i64, i64, c_float, vp]
    L.dfot_dwconv3x3_glu_bf16.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, vp]
    L.dfot_rmsnorm_affine.argtypes = [vp, i64, vp, vp, c_float, vp, i, vp, vp, i64, i64, vp]
    for name in SYMBOLS:
        if name not in ("dfot_last_error", "dfot_launch_count"):
            getattr(L, name).restype = c_int
    if L.dfot_abi_version() != 1:
        raise RuntimeError("dfot_b200: ABI version mismatch between Python binding and libdfot_b200.so")
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().dfot_last_error().decode(errors="replace")
        raise RuntimeError(f"dfot_b200 {what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().dfot_launch_count())
