"""ORACLE (test infrastructure): config trees of the golden / parity cases.

``algorithm_cfg`` builds the full ``cfg.algorithm`` tree the reference reads
eagerly in BaseVideoAlgo.__init__ (base_pytorch_video_algo.py:37-88) with the
defaults of configurations/algorithm/dfot_video.yaml; keyword overrides use
dotted paths ("diffusion.sampling_timesteps").
"""
import copy
from typing import Any, Dict


def _set(d: dict, path: str, value: Any) -> None:
    keys = path.split(".")
    for k in keys[:-1]:
        d = d.setdefault(k, {})
    d[keys[-1]] = value


def algorithm_cfg(**overrides) -> Dict[str, Any]:
    overrides = dict(overrides)
    backbone_override = overrides.pop("backbone", None)
    cfg = dict(
        debug=False, lr=1e-4,
        external_cond_type=None, external_cond_num_classes=None, external_cond_dim=0,
        external_cond_stack=False, external_cond_processing=None,
        backbone=dict(name="dit3d", variant="full", pos_emb_type="rope_3d", patch_size=2, hidden_size=384,
                      depth=12, num_heads=6, mlp_ratio=4.0, use_gradient_checkpointing=False),
        x_shape=[4, 16, 16], max_frames=8, n_frames=8, frame_skip=1, context_frames=1,
        latent=dict(enabled=False, type="pre_sample", suffix=None, downsampling_factor=[1, 8], shape=None,
                    num_channels=4),
        data_mean=[[[0.0]]] * 4, data_std=[[[1.0]]] * 4,
        compile=False, weight_decay=1e-3, optimizer_beta=[0.9, 0.99],
        lr_scheduler=dict(name="constant_with_warmup", num_warmup_steps=5000),
        noise_level="random_independent", uniform_future=dict(enabled=False),
        fixed_context=dict(enabled=False, indices=None, dropout=0),
        variable_context=dict(enabled=False, prob=0.25, dropout=0.3),
        chunk_size=-1, scheduling_matrix="full_sequence", replacement="noisy_scale",
        refinement_sampling=dict(enabled=False, goback_length=20, n_goback=5),
        save_attn_map=dict(enabled=False, attn_map_dir=""),
        diffusion=dict(is_continuous=False, timesteps=1000, beta_schedule="cosine",
                       schedule_fn_kwargs=dict(shift=1.0), use_causal_mask=False, clip_noise=20.0,
                       objective="pred_v",
                       loss_weighting=dict(strategy="fused_min_snr", snr_clip=5.0, cum_snr_decay=0.9),
                       sampling_timesteps=50, ddim_sampling_eta=0.0, reconstruction_guidance=0.0),
        vae=dict(pretrained_path=None, pretrained_kwargs={}, use_fp16=False, batch_size=2),
        checkpoint=dict(reset_optimizer=False, strict=True),
        tasks=dict(
            prediction=dict(enabled=True, history_guidance=dict(name="conditional", visualize=False),
                            keyframe_density=None, sliding_context_len=None),
            interpolation=dict(enabled=False, history_guidance=dict(name="conditional", visualize=False),
                               max_batch_size=None)),
        logging=dict(deterministic=0, loss_freq=100, grad_norm_freq=100, max_num_videos=8,
                     n_metrics_frames=None, metrics=[], metrics_batch_size=16, sanity_generation=False,
                     raw_dir=None),
    )
    if backbone_override is not None:
        cfg["backbone"] = copy.deepcopy(backbone_override)
    for k, v in overrides.items():
        _set(cfg, k.replace("__", "."), copy.deepcopy(v))
    return cfg


def continuous_overrides() -> Dict[str, Any]:
    """configurations/shortcut/diffusion/continuous.yaml + the dataset_experiment defaults it is used with."""
    return {
        "diffusion.is_continuous": True, "diffusion.precond_scale": 0.125,
        "diffusion.beta_schedule": "cosine_simple_diffusion",
        "diffusion.schedule_fn_kwargs": dict(shifted=0.125, interpolated=False),
        "diffusion.training_schedule": dict(name="cosine", shift=0.125),
        "diffusion.loss_weighting": dict(strategy="sigmoid", sigmoid_bias=-1.0),
        "backbone.use_fourier_noise_embedding": True,
    }


def _small(**kw):
    """golden-sized DiT3D: D=64, 1 head (d=64), depth 2, MLP x2, latents [4,8,8], patch 2 → P=16; 4 tokens/window."""
    base = {
        "backbone.hidden_size": 64, "backbone.depth": 2, "backbone.num_heads": 1,
        "backbone.spatial_mlp_ratio": 2.0, "x_shape": [4, 8, 8], "max_frames": 4, "n_frames": 4,
        "context_frames": 1, "diffusion.sampling_timesteps": 5,
    }
    base.update(kw)
    return algorithm_cfg(**base)


def golden_cases() -> Dict[str, Dict[str, Any]]:
    """name -> dict(cfg, batch, n_frames, cond_dim, weights)  (weights = which state-dict the case uses)"""
    act = {"external_cond_type": "action", "external_cond_dim": 3, "external_cond_processing": "mask_first",
           "backbone.external_cond_dropout": 0.1}
    cases = {
        "vanilla": dict(cfg=_small(**{"tasks.prediction.history_guidance":
                                      dict(name="vanilla", guidance_scale=4.0, visualize=False)}),
                        batch=2, weights="plain"),
        "conditional_nomlp": dict(cfg=_small(**{"backbone.spatial_mlp_ratio": None}), batch=2, weights="nomlp"),
        "stabilized_sliding": dict(cfg=_small(**{
            "n_frames": 7, "tasks.prediction.sliding_context_len": 2,
            "tasks.prediction.history_guidance": dict(name="stabilized_vanilla", guidance_scale=2.0,
                                                      stabilization_level=0.02, visualize=False)}),
            batch=1, weights="plain"),
        "pyramid_conditional": dict(cfg=_small(**{"scheduling_matrix": "autoregressive"}), batch=1, weights="plain"),
        "full_sequence_fractional": dict(cfg=_small(**{
            "noise_level": "random_uniform", "context_frames": 2,
            "tasks.prediction.history_guidance": dict(name="fractional", guidance_scale=3.0, freq_scale=0.3,
                                                      visualize=False)}), batch=1, weights="plain"),
        "temporal": dict(cfg=_small(**{
            "context_frames": 2,
            "tasks.prediction.history_guidance": dict(name="temporal", hist_subsequences=[[0], [1]],
                                                      hist_weights=[1.5, 1.5], gen_segments=[[0], [1], [0, 1]],
                                                      visualize=False)}), batch=1, weights="plain"),
        "continuous_action": dict(cfg=_small(**{**continuous_overrides(), **act,
                                                "tasks.prediction.history_guidance":
                                                dict(name="vanilla", guidance_scale=2.5, visualize=False)}),
                                  batch=2, weights="action"),
        "keyframes_interp": dict(cfg=_small(**{
            **continuous_overrides(), **act, "n_frames": 9,
            "tasks.prediction.keyframe_density": 0.5, "tasks.prediction.sliding_context_len": 1,
            "tasks.interpolation.enabled": False,
            "tasks.interpolation.history_guidance": dict(name="vanilla", guidance_scale=1.5, visualize=False),
            "tasks.interpolation.max_batch_size": 2}), batch=1, weights="action"),
    }
    pose = {
        **continuous_overrides(), "external_cond_type": "action", "external_cond_dim": 16,
        "camera_pose_conditioning": dict(normalize_by="first", bound=None, type="ray_encoding"),
        "backbone": dict(name="u_vit3d_pose", channels=[32, 32, 64, 128], emb_channels=64, patch_size=2,
                         block_types=["ResBlock", "ResBlock", "TransformerBlock", "TransformerBlock"],
                         block_dropouts=[0.0, 0.0, 0.0, 0.0], num_updown_blocks=[1, 1, 2], num_mid_blocks=2,
                         num_heads=1, pos_emb_type="rope", use_checkpointing=[False] * 4,
                         conditioning=dict(dim=None), external_cond_dropout=0.1, use_fourier_noise_embedding=True),
        "x_shape": [3, 32, 32], "max_frames": 4, "n_frames": 4, "context_frames": 1,
        "data_mean": [[[0.5]]] * 3, "data_std": [[[0.5]]] * 3, "diffusion.sampling_timesteps": 3,
    }
    cases["uvit_pose_vanilla"] = dict(cfg=algorithm_cfg(**{**pose, "tasks.prediction.history_guidance":
                                                          dict(name="vanilla", guidance_scale=2.0, visualize=False)}),
                                      batch=1, weights="uvit_pose", algo="dfot_video_pose")
    cases["uvit_pose_stabilized_interp"] = dict(cfg=algorithm_cfg(**{
        **pose, "n_frames": 9, "tasks.prediction.keyframe_density": 0.5, "tasks.prediction.sliding_context_len": 1,
        "tasks.prediction.history_guidance": dict(name="stabilized_vanilla", guidance_scale=2.0,
                                                  stabilization_level=0.02, visualize=False),
        "tasks.interpolation.history_guidance": dict(name="vanilla", guidance_scale=1.5, visualize=False),
        "tasks.interpolation.max_batch_size": 2}), batch=1, weights="uvit_pose", algo="dfot_video_pose")
    return cases


def synthetic_poses(batch: int, n_frames: int):
    """(B, T, 16): intrinsics (fx, fy, px, py) = (0.5, 0.9, 0.5, 0.5) + row-major [R | t] of a smooth yaw/pitch/
    translation trajectory (valid rotations), as in SURVEY.md §8c."""
    import math

    import torch
    out = []
    for b in range(batch):
        rows = []
        for t in range(n_frames):
            yaw, pitch = 0.05 * t + 0.1 * b, 0.02 * t
            cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
            R = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ torch.tensor([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
            tv = torch.tensor([0.1 * t, 0.02 * t * (b + 1), 0.05 * t + 0.3])
            rows.append(torch.cat([torch.tensor([0.5, 0.9, 0.5, 0.5]), torch.cat([R, tv[:, None]], 1).flatten()]))
        out.append(torch.stack(rows))
    return torch.stack(out).float()


# matrix-attention configurations beyond the golden rollouts: (variant, matrix_block, embed_col_dim, num_col_heads,
# num_row_heads, embed_row_dim, use_temporal_rope, flatten_matrix_rope, matrix_multi_token, use_bias, fixed_u)
MATRIX_COMBOS = [
    ("full_matrix_attention", "matrix_self", 2, 2, 1, 64, True, False, False, False, None),
    ("factorized_matrix_attention", "matrix_cross", 1, 1, 1, 64, False, False, False, False, None),
    ("full_matrix_attention", "matrix", 4, 2, 1, 64, True, True, False, True, None),         # grouped + flatten + bias rows
    ("factorized_matrix_attention", "matrix", 4, 1, 2, 64, False, False, False, True, None),  # 4 rows x 32 = 128 wide, no RoPE
    ("full_matrix_attention", "matrix", 3, 3, 2, 128, True, False, True, True, None),        # odd column count, two row heads
    ("full_matrix_attention", "matrix_cross", 4, 2, 1, 64, True, False, True, False, None),
    ("factorized_matrix_attention", "matrix_self", 2, 1, 1, 64, True, True, False, False, None),
    ("full_matrix_attention", "matrix", 16, 16, 1, 64, True, False, False, False, "identity"),
]


def matrix_combo_cfg(combo) -> dict:
    """Backbone config of one MATRIX_COMBOS entry at golden size (8x8 latents, patch 2, 4 frames, depth 2)."""
    variant, block, ecd, nch, nrh, erd, rope, flat, multi, bias, fixed_u = combo
    return _small(**{"backbone.variant": variant, "backbone.pos_emb_type": "sinusoidal_2d", "backbone.use_temporal_rope": rope,
                     "backbone.hidden_size": erd, "backbone.embed_col_dim": ecd, "backbone.embed_row_dim": erd,
                     "backbone.num_col_heads": nch, "backbone.num_row_heads": nrh, "backbone.num_heads": erd // 64,
                     "backbone.mlp_ratio": 2.0, "backbone.spatial_mlp_ratio": 2.0, "backbone.use_bias": bias,
                     "backbone.matrix_block": block, "backbone.flatten_matrix_rope": flat,
                     "backbone.matrix_multi_token": multi, "backbone.fixed_u": fixed_u})["backbone"]
