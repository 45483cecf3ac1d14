#!/usr/bin/env python
"""bench.py — DFoT denoising-sampling throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU baseline arm (reference algorithm on host cores)

Workloads (BASELINE.json `configs`):
  re10k (default; the configuration the metric "RE10K-shaped DFoT + HG" is quoted on, configs[2]): dfot_video_pose,
        UViT3DPose (channels 128/256/576/1152, blocks 3/3/6 + 20 mid, 9 heads, emb 1024), 8 frames 256x256 pixels,
        1 context frame, vanilla history guidance scale 4.0 (2 branches), 50 DDIM steps, batch 4 per GPU.
  k600  (configs[1]): K600-shaped DiT3D-XL latent sampling — latents [16,16,16], 17 frames = 5 tokens (2 context),
        50 DDIM steps, conditional history guidance (nfe=1), batch 8 per GPU.
A bench "step" is one full sampling pass over one batch (50 denoising steps).  Synthetic inputs, random-init weights
(zero-initialised output layers re-drawn N(0,0.02) so the network is not identically 0).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# --------------------------------------------------------------------------------------- workload definition
def k600_cfg(sampling_timesteps=50, spatial_mlp_ratio=4.0, depth=28, hidden=1152, heads=16):
    """`dataset=kinetics_600 algorithm=dfot_video @DiT/XL` resolved by hand (kinetics_600.yaml,
    kinetics_600_video_generation.yaml, shortcut/DiT/XL.yaml, dfot_video.yaml).  spatial_mlp_ratio=4.0 gives the
    MLP blocks of the published DiT-XL (the fork's dit3d.yaml leaves it unset → no MLP)."""
    mean = [[[0.0]]] * 16
    std = [[[1.0]]] * 16
    return dict(
        debug=False, lr=1e-4, external_cond_type=None, external_cond_num_classes=None, external_cond_dim=0,
        external_cond_stack=False, external_cond_processing=None,
        backbone=dict(name="dit3d", variant="full", pos_emb_type="rope_3d", patch_size=1, hidden_size=hidden,
                      depth=depth, num_heads=heads, mlp_ratio=4.0, spatial_mlp_ratio=spatial_mlp_ratio,
                      use_gradient_checkpointing=False),
        x_shape=[3, 128, 128], max_frames=17, n_frames=17, frame_skip=1, context_frames=5,
        latent=dict(enabled=True, type="pre_sample", suffix=None, downsampling_factor=[4, 8], shape=None,
                    num_channels=16),
        data_mean=mean, data_std=std, compile=False, weight_decay=0, optimizer_beta=[0.9, 0.99],
        lr_scheduler=dict(name="constant_with_warmup", num_warmup_steps=10000), noise_level="random_independent",
        uniform_future=dict(enabled=False), fixed_context=dict(enabled=False, indices=None, dropout=0),
        variable_context=dict(enabled=False, prob=0.25, dropout=0.3), chunk_size=-1,
        scheduling_matrix="full_sequence", replacement="noisy_scale",
        refinement_sampling=dict(enabled=False, goback_length=20, n_goback=5),
        save_attn_map=dict(enabled=False, attn_map_dir=""),
        diffusion=dict(is_continuous=False, timesteps=1000, beta_schedule="cosine", schedule_fn_kwargs=dict(shift=1.0),
                       use_causal_mask=False, clip_noise=20.0, objective="pred_v",
                       loss_weighting=dict(strategy="fused_min_snr", snr_clip=5.0, cum_snr_decay=0.96),
                       sampling_timesteps=sampling_timesteps, ddim_sampling_eta=0.0, reconstruction_guidance=0.0),
        vae=dict(pretrained_path=None, pretrained_kwargs={}, use_fp16=False, batch_size=16),
        checkpoint=dict(reset_optimizer=False, strict=True),
        tasks=dict(prediction=dict(enabled=True, history_guidance=dict(name="conditional", visualize=False),
                                   keyframe_density=None, sliding_context_len=None),
                   interpolation=dict(enabled=False, history_guidance=dict(name="conditional", visualize=False),
                                      max_batch_size=None)),
        logging=dict(deterministic=0, loss_freq=100, grad_norm_freq=100, max_num_videos=8, n_metrics_frames=None,
                     metrics=[], metrics_batch_size=16, sanity_generation=False, raw_dir=None))


def dmlab_cfg(sampling_timesteps=50, frames=36, guidance_scale=None):
    """`dataset=dmlab algorithm=dfot_video @diffusion/continuous @DiT/B dataset.max_frames=T` (BASELINE config[4]) resolved
    by hand: dmlab.yaml (latents 32 ch, 64 px / 8 -> [32, 8, 8]; action dim 3 stacked over frame_skip 1, mask_first),
    dmlab_video_generation.yaml (patch 2, external_cond_dropout 0.1, sigmoid weighting, shifted cosine 0.125),
    shortcut/DiT/B.yaml (768 x 12, 12 heads), shortcut/diffusion/continuous.yaml; context_length 4 (base_video.yaml)."""
    cfg = k600_cfg(sampling_timesteps, 4.0, depth=12, hidden=768, heads=12)
    cfg.update(external_cond_type="action", external_cond_dim=3, external_cond_stack=True,
               external_cond_processing="mask_first", x_shape=[3, 64, 64], max_frames=frames, n_frames=frames,
               context_frames=4, data_mean=[[[0.0]]] * 32, data_std=[[[3.46140533056]]] * 32,
               latent=dict(enabled=True, type="pre_sample", suffix=None, downsampling_factor=[1, 8], shape=None,
                           num_channels=32))
    cfg["backbone"].update(patch_size=2, external_cond_dropout=0.1, use_fourier_noise_embedding=True)
    cfg["diffusion"].update(is_continuous=True, precond_scale=0.125, beta_schedule="cosine_simple_diffusion",
                            schedule_fn_kwargs=dict(shifted=0.125, interpolated=False),
                            training_schedule=dict(name="cosine", shift=0.125),
                            loss_weighting=dict(strategy="sigmoid", sigmoid_bias=-1.0))
    if guidance_scale:
        cfg["tasks"]["prediction"]["history_guidance"] = dict(name="vanilla", guidance_scale=guidance_scale,
                                                              visualize=False)
    return cfg


# configurations/algorithm/dc_ae_preprocessor.yaml (the `vae` node of dmlab_video_generation.yaml), decode side
DCAE_DMLAB_CFG = dict(in_channels=3, latent_channels=32, attention_head_dim=32, scaling_factor=0.2889,
                      decoder_block_types=["ResBlock", "ResBlock", "ResBlock", "EfficientViTBlock"],
                      decoder_block_out_channels=[128, 256, 512, 512], decoder_layers_per_block=[0, 5, 10, 2],
                      decoder_norm_types=["batch_norm", "batch_norm", "batch_norm", "rms_norm"],
                      decoder_act_fns=["relu", "relu", "relu", "silu"], decoder_qkv_multiscales=[[], [], [], []],
                      upsample_block_type="pixel_shuffle")


def re10k_cfg(sampling_timesteps=50, guidance_scale=4.0):
    """`dataset=realestate10k_mini algorithm=dfot_video_pose @diffusion/continuous dataset.context_length=1
    dataset.n_frames=8 ...history_guidance.name=vanilla +guidance_scale=4.0` (README.md:74) resolved by hand from
    dfot_video_pose.yaml, backbone/u_vit3d_pose.yaml, dataset_experiment/realestate10k_video_generation.yaml,
    dataset/realestate10k.yaml, shortcut/diffusion/continuous.yaml."""
    cfg = k600_cfg(sampling_timesteps)
    cfg.update(
        external_cond_type="action", external_cond_dim=16,
        camera_pose_conditioning=dict(normalize_by="first", bound=None, type="ray_encoding"),
        backbone=dict(name="u_vit3d_pose", channels=[128, 256, 576, 1152], emb_channels=1024, patch_size=2,
                      block_types=["ResBlock", "ResBlock", "TransformerBlock", "TransformerBlock"],
                      block_dropouts=[0.0, 0.0, 0.1, 0.1], num_updown_blocks=[3, 3, 6], num_mid_blocks=20, num_heads=9,
                      pos_emb_type="rope", use_checkpointing=[False, False, False, True], conditioning=dict(dim=None),
                      external_cond_dropout=0.1, use_fourier_noise_embedding=True),
        x_shape=[3, 256, 256], max_frames=8, n_frames=8, frame_skip=20, context_frames=1,
        latent=dict(enabled=False, type="pre_sample", suffix=None, downsampling_factor=[1, 1], shape=None,
                    num_channels=3),
        data_mean=[[[0.577]], [[0.517]], [[0.461]]], data_std=[[[0.249]], [[0.249]], [[0.268]]])
    cfg["diffusion"].update(is_continuous=True, precond_scale=0.125, beta_schedule="cosine_simple_diffusion",
                            schedule_fn_kwargs=dict(shifted=0.125, interpolated=False),
                            training_schedule=dict(name="cosine", shift=0.125),
                            loss_weighting=dict(strategy="sigmoid", sigmoid_bias=-1.0))
    cfg["tasks"]["prediction"]["history_guidance"] = dict(name="vanilla", guidance_scale=guidance_scale, visualize=False)
    return cfg


class Workload:
    """Shapes and counters of one BASELINE.json configuration."""

    def __init__(self, name, args):
        self.name = name
        if name == "k600":
            self.cfg = k600_cfg(args.sampling_steps, None if args.no_mlp else 4.0)
            self.batch = args.batch or 8
            self.n_tokens, self.ctx_tokens, self.gen_frames, self.nfe = 5, 2, 17 - 5, 1
            self.x_shape = [16, 16, 16]
            self.text = (f"K600-shaped DFoT DiT3D-XL (28x1152, 16 heads d=72, patch 1, "
                         f"{'no MLP (fork default)' if args.no_mlp else 'MLP x4'}) latent sampling: latents 16x16x16, "
                         f"17 frames = 5 tokens (2 context), {args.sampling_steps} DDIM steps, conditional HG (nfe=1), "
                         f"batch {self.batch}/GPU")
            self.l2 = "per-step activations (~0.5 GB) and weights (1.3 GB bf16) exceed the 126 MB L2; no flush needed"
        elif name == "dmlab":
            T = args.frames or 36
            hg = args.guidance or 0.0
            self.cfg = dmlab_cfg(args.sampling_steps, T, hg if hg > 1.0 else None)
            self.batch = args.batch or 16
            self.n_tokens, self.ctx_tokens, self.gen_frames, self.nfe = T, 4, T - 4, 2 if hg > 1.0 else 1
            self.x_shape = [32, 8, 8]
            self.text = (f"DMLab-shaped long-context DFoT DiT3D-B (12x768, 12 heads d=64, patch 2, MLP x4), continuous "
                         f"diffusion, latents 32x8x8, {T} frames = {T * 16} tokens (4 context), action conditioning "
                         f"(mask_first), {args.sampling_steps} DDIM steps, "
                         f"{'vanilla HG %.1f (nfe=2)' % hg if hg > 1.0 else 'conditional HG (nfe=1)'}, batch {self.batch}/GPU")
            self.l2 = ("small model (170 MB of bf16 weights): activations of a step fit the 126 MB L2 at small batch; "
                       "256 MiB of HBM are overwritten between timed passes to flush it")
        elif name == "re10k_long":
            # README.md:69 "Single Image to Long Video (200 Frames)" = BASELINE config[3]
            self.cfg = re10k_cfg(args.sampling_steps)
            self.cfg.update(n_frames=200, frame_skip=1)
            self.cfg["tasks"]["prediction"].update(
                history_guidance=dict(name="stabilized_vanilla", guidance_scale=4.0, stabilization_level=0.02,
                                      visualize=False), keyframe_density=0.0625)
            self.cfg["tasks"]["interpolation"].update(
                history_guidance=dict(name="vanilla", guidance_scale=1.5, visualize=False), max_batch_size=4)
            self.batch = args.batch or 1
            self.n_tokens, self.ctx_tokens, self.gen_frames, self.nfe = 200, 1, 200 - 1, 2
            self.x_shape = [3, 256, 256]
            self.text = (f"RE10K single-image-to-200-frame rollout (dfot_video_pose, UViT3DPose as in the short workload): "
                         f"12 keyframes by two sliding windows under stabilized_vanilla(4.0, 0.02), then two rounds of "
                         f"vanilla(1.5) keyframe interpolation (11 + 35 chunks in batches of 4), {args.sampling_steps} "
                         f"DDIM steps, {self.batch} sample(s) in total (strong scaling: chunk batches dealt over the "
                         f"sample axis of the mesh, branches split inside pairs)")
            self.l2 = ("per-forward activations (>1 GB per row), pose modulation cache and weights (1.1 GB bf16) exceed "
                       "the 126 MB L2; no flush needed")
        else:
            self.cfg = re10k_cfg(args.sampling_steps)
            self.batch = args.batch or 4
            self.n_tokens, self.ctx_tokens, self.gen_frames, self.nfe = 8, 1, 8 - 1, 2
            self.x_shape = [3, 256, 256]
            self.text = (f"RE10K-shaped dfot_video_pose single-image-to-short: UViT3DPose (128/256/576/1152 ch, blocks "
                         f"3/3/6 + 20 mid, 9 heads, emb 1024), 8 frames 256x256 (1 context), vanilla history guidance "
                         f"4.0 (2 branches), {args.sampling_steps} DDIM steps, batch {self.batch}/GPU")
            self.l2 = ("per-forward activations (>1 GB per row), pose modulation cache (~1 GB per sample) and weights "
                       "(1.1 GB bf16) exceed the 126 MB L2; no flush needed")

    def inputs(self, rank):
        import torch
        if self.name == "re10k_long":
            rank = 0                                   # strong scaling: every rank holds the same sample(s)
        g = torch.Generator().manual_seed(123 + rank)
        if self.name == "k600":
            return torch.randn((self.batch, self.n_tokens, *self.x_shape), generator=g), None
        if self.name == "dmlab":
            return (torch.randn((self.batch, self.n_tokens, *self.x_shape), generator=g),
                    torch.randn((self.batch, self.n_tokens, 3), generator=g))
        xs = torch.rand((self.batch, self.n_tokens, *self.x_shape), generator=g)
        return xs, synthetic_poses(self.batch, self.n_tokens)

    def forward_row_gflop(self):
        """Algorithmic GFLOP (2*MAC of GEMM / conv / attention) the sampler must execute per backbone forward-row.
        DiT: per block per token 8D^2 + 4ND (+ 4rD^2), adaLN per frame.  U-ViT: ResBlock 36*C^2 per
        pixel, TransformerBlock 24*C^2 + 4*N*C per token, resampling convs 18*Cin*Cout per output pixel; the
        camera-pose PatchEmbed and pose part of every emb_layer are constant per window and NOT counted (the
        reference executes them every step: 6626 GFLOP/row there)."""
        b = self.cfg["backbone"]
        if self.name in ("k600", "dmlab"):
            D, depth, r = b["hidden_size"], b["depth"], b.get("spatial_mlp_ratio") or 0
            N = self.n_tokens * (self.x_shape[1] // b["patch_size"]) ** 2
            return depth * N * (8 * D * D + 4 * N * D + 4 * r * D * D) / 1e9
        T, ch, L = self.cfg["max_frames"], b["channels"], len(b["channels"])
        res = [self.x_shape[1] // b["patch_size"] // 2 ** i for i in range(L)]
        nblk = [2 * n for n in b["num_updown_blocks"]] + [b["num_mid_blocks"]]
        total = 0.0
        for i in range(L):
            px = T * res[i] ** 2
            if b["block_types"][i] == "ResBlock":
                total += nblk[i] * px * 36 * ch[i] ** 2
            else:
                total += nblk[i] * px * (24 * ch[i] ** 2 + 4 * px * ch[i])
            if i + 1 < L:
                total += 2 * (T * res[i + 1] ** 2) * 18 * ch[i] * ch[i + 1]
        total += 2 * T * res[0] ** 2 * 2 * ch[0] * 12
        return total / 1e9


def synthetic_poses(batch, n_frames):
    """(B, T, 16) = intrinsics (fx, fy, px, py) = (0.5, 0.9, 0.5, 0.5) + row-major [R | t] of a smooth
    yaw / pitch / translation trajectory (valid rotations) — SURVEY.md §8c synthetic inputs."""
    import torch
    out = torch.zeros((batch, n_frames, 16))
    for b in range(batch):
        for t in range(n_frames):
            yaw, pitch = 0.05 * t + 0.1 * b, 0.02 * t
            cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
            R = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ torch.tensor([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
            tv = torch.tensor([0.1 * t, 0.02 * t * (b + 1), 0.05 * t + 0.3])
            out[b, t] = torch.cat([torch.tensor([0.5, 0.9, 0.5, 0.5]), torch.cat([R, tv[:, None]], 1).flatten()])
    return out


def make_weights(cfg, seed=0):
    import torch
    from dfot_b200.algorithms.dfot import DFoTVideo, DFoTVideoPose
    torch.manual_seed(seed)
    algo = (DFoTVideoPose if cfg["backbone"]["name"] == "u_vit3d_pose" else DFoTVideo)(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for _, p in algo.named_parameters():
            if bool((p == 0).all()):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return algo


# --------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------- CPU baseline
def _bounded_cfg(wl, n_steps):
    cfg = json.loads(json.dumps(wl.cfg))
    cfg["diffusion"]["sampling_timesteps"] = n_steps
    return cfg


def _baseline_record(wl, sampling_steps, kind, what, cores, n_steps, rows, dt):
    nfe_s = rows / dt
    return dict(value=nfe_s * wl.gen_frames / (sampling_steps * wl.nfe), unit="generated_frames/s", cores=cores,
                kind=kind, nfe_per_sec=nfe_s,
                sample=f"{what}, torch fp32 CPU, {cores} threads: batch 1 x {n_steps} DDIM step(s) x {wl.nfe} branch(es) x "
                       f"{rows // (n_steps * wl.nfe)} pass(es) = {rows} forward-rows of the full-size backbone in {dt:.1f}s; "
                       f"frames/s derived as NFE/s*{wl.gen_frames}/({sampling_steps}*{wl.nfe})")


def reference_available():
    from oracle import ref_shim
    return ref_shim.available()


def cpu_baseline(wl, sampling_steps, seconds_budget=25.0, prefer_reference=True):
    """The reference's own CPU implementation of the path on a bounded sample of the same workload: batch 1 and 1-2
    sampling steps of the full-size backbone, all host threads.  `kind = "reference"`: the UNMODIFIED reference files
    (oracle/_ref, copied from /root/reference by oracle/build_ref.py) driven through its public `_predict_videos`;
    `kind = "port"` (only when that copy is absent): the oracle restatement, pinned to the reference <= 1e-5.
    NFE/s is step-count independent; frames/s = NFE/s * generated frames / (sampling steps * nfe)."""
    import torch
    if wl.name == "re10k_long":
        # same backbone and per-row cost as the 8-frame workload: time that bounded sample and convert with the long
        # rollout's row count (96 forward-rows per DDIM step per sample: 2 + 2 keyframe-window rows, 22 + 70 chunk rows)
        short = Workload("re10k", type("A", (), dict(sampling_steps=sampling_steps, no_mlp=False, batch=1))())
        cb = cpu_baseline(short, sampling_steps, seconds_budget, prefer_reference)
        cb["value"] = cb["nfe_per_sec"] * wl.gen_frames / (sampling_steps * 96)
        cb["sample"] += f"; long rollout: frames/s = NFE/s*{wl.gen_frames}/({sampling_steps}*96 rows per step)"
        return cb
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_steps = 2 if wl.name in ("k600", "dmlab") else 1
    cfg = _bounded_cfg(wl, n_steps)
    xs, conds = wl.inputs(0)
    xs, conds = xs[:1], None if conds is None else conds[:1]
    if prefer_reference and reference_available():
        from oracle import ref_shim
        ref_shim.install()
        from algorithms.dfot.dfot_video import DFoTVideo as RefVideo
        from algorithms.dfot.dfot_video_pose import DFoTVideoPose as RefVideoPose
        if cfg["latent"]["enabled"] and cfg["latent"]["downsampling_factor"][0] > 1:
            cfg["latent"]["type"] = "online"     # kinetics_600.yaml:9 (the reference asserts it for VideoVAE latents)
        torch.manual_seed(0)
        ref = (RefVideoPose if cfg["backbone"]["name"] == "u_vit3d_pose" else RefVideo)(ref_shim.to_dc(cfg)).eval()
        ref_shim.rerandomize_zero_params(ref, 1)
        if wl.name.startswith("re10k"):
            xs = ref._normalize_x(xs)
        run = lambda: ref._predict_videos(xs.clone(), n_context_tokens=wl.ctx_tokens, conditions=conds)
        kind = "reference"
        what = (f"the reference itself (unmodified files, {os.path.relpath(ref_shim.REF, ROOT) if ref_shim.REF.startswith(ROOT) else ref_shim.REF}; "
                "public API `_predict_videos`)")
    else:
        from oracle.sampler import SamplerOracle
        algo = make_weights(cfg, 0)
        weights = {k[len("diffusion_model.model."):]: v.detach() for k, v in algo.state_dict().items()
                   if k.startswith("diffusion_model.model.")}
        probe = SamplerOracle(cfg, None)
        if wl.name in ("k600", "dmlab"):
            from oracle.dit3d import DiT3DOracle
            model = DiT3DOracle(cfg["backbone"], probe.x_shape, probe.max_tokens, weights,
                                external_cond_dim=probe.external_cond_dim)
        else:
            from oracle.uvit3d_pose import UViT3DPoseOracle
            model = UViT3DPoseOracle(cfg["backbone"], probe.x_shape, probe.max_tokens, weights)
        del algo
        oracle = SamplerOracle(cfg, model)
        run = lambda: oracle.predict_videos(xs.clone(), wl.ctx_tokens, conds)
        kind, what = "port", "oracle port of the reference algorithm"
    torch.manual_seed(123)
    t0 = time.perf_counter()
    rows = 0
    with torch.no_grad():
        while True:
            run()
            rows += n_steps * wl.nfe
            if time.perf_counter() - t0 > seconds_budget * 0.5 or rows >= 8:
                break
    dt = time.perf_counter() - t0
    return _baseline_record(wl, sampling_steps, kind, what, cores, n_steps, rows, dt)


def parity_check(wl, dev):
    """Outside the timed region: ONE sampling step of the benchmarked full-size model at batch 1 on the GPU (this repo's
    kernels) and on the CPU oracle (torch fp32 restatement of the reference, pinned <= 1e-5 to the executed reference) with
    the same weights, inputs and noise; north_star gates: denoiser output max-abs <= 2e-2, PSNR >= 40 dB."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import NoiseBank, build_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = _bounded_cfg(wl, 1)
    algo = make_weights(cfg, 0)
    weights = {k[len("diffusion_model.model."):]: v.detach().clone() for k, v in algo.state_dict().items()
               if k.startswith("diffusion_model.model.")}
    xs, conds = wl.inputs(0)
    xs, conds = xs[:1], None if conds is None else conds[:1]
    if wl.name.startswith("re10k"):
        xs = algo._normalize_x(xs)
    bank = NoiseBank(51)
    oracle, _ = build_oracle(json.loads(json.dumps(cfg)), weights, randn=bank.randn, randn_like=bank.randn_like)
    oracle.trace = []
    with torch.no_grad():
        ref = oracle.predict_videos(xs.clone(), wl.ctx_tokens, conds)
    bank2 = NoiseBank(51)
    algo = algo.to(dev).eval()
    algo.diffusion_model.noise_source = lambda shape, device: bank2.randn(shape).to(device)
    algo.trace = []
    out = algo._predict_videos(xs.to(dev), wl.ctx_tokens, None if conds is None else conds.to(dev)).cpu()
    err = max((t["model_out"].cpu() - o["model_out"]).abs().max().item() for t, o in zip(algo.trace, oracle.trace))
    exact = all((t["levels_from"] == o["levels_from"].numpy()).all() and (t["levels_to"] == o["levels_to"].numpy()).all()
                for t, o in zip(algo.trace, oracle.trace))
    n = wl.ctx_tokens
    rng = (ref[:, n:].max() - ref[:, n:].min()).item()
    mse = ((out[:, n:].double() - ref[:, n:].double()) ** 2).mean().item()
    psnr = 10 * math.log10(rng * rng / max(mse, 1e-30))
    return dict(max_abs=err, psnr=psnr, levels_bit_exact=bool(exact), gate=dict(max_abs=2e-2, psnr=40.0),
                ok=bool(err <= 2e-2 and psnr >= 40.0 and exact),
                sample=f"batch 1 x 1 DDIM step x {wl.nfe} branch row(s) of the benchmarked full-size model vs the CPU oracle "
                       "(same weights / inputs / noise), per-step denoiser output and the step's sample")


# --------------------------------------------------------------------------------------- communicating multi-GPU paths
def strong_section(args, wl, algo, dev, world):
    """N >= 2, after the weak-scaling measurement: the two paths of SURVEY.md §8e that DO communicate inside the sampling
    loop, each timed against the same work on ONE GPU in the same process group (every rank runs the single-GPU version
    at the same time, so both sides see the same box under the same load) and checked for equality on the device:
      branch_split : one sample per pair of GPUs, the two history-guidance branch rows of every step split over the pair,
                     one NCCL all-gather of the backbone output per step, identical fused K4 step on both members;
      long_rollout : BASELINE config[3] (one 200-frame video: keyframe windows + two interpolation rounds) at
                     `--strong-sampling-steps` DDIM steps over the dp x br mesh (NFE/s does not depend on the step count).
    Device-timed (CUDA events, barrier + synchronize on both sides), max over ranks."""
    import torch
    import torch.distributed as dist
    from dfot_b200 import distributed as D
    from dfot_b200.algorithms.dfot import DFoTVideoPose

    def timed(fn, reps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), r

    def max_over_ranks(v):
        t = torch.tensor([float(v)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    br = 2 if world % 2 == 0 else 1
    mesh = D.build_mesh(br=br)
    res = dict(mesh=f"dp {mesh.dp} x br {mesh.br}")

    def with_mesh(a, m, seed, fn):
        a.mesh = m
        torch.manual_seed(seed)
        try:
            return fn()
        finally:
            a.mesh = None

    # ---- (i) history-guidance branches split inside a pair
    if br == 2:
        xs_h, conds_h = wl.inputs(mesh.dp_index)
        xs1 = algo._normalize_x(xs_h[:1].to(dev))
        c1 = conds_h[:1].to(dev)
        seed = 777 + mesh.dp_index
        single = lambda: with_mesh(algo, None, seed, lambda: algo._predict_videos(xs1, wl.ctx_tokens, c1))
        split = lambda: with_mesh(algo, mesh, seed, lambda: algo._predict_videos(xs1, wl.ctx_tokens, c1))
        for _ in range(2):      # eager + graph capture, then a replay, for the 2-row and the 1-row forward
            single()
            split()
        t1, o1 = timed(single, 2)
        t2, o2 = timed(split, 2)
        local = torch.randn((1, wl.n_tokens, *wl.x_shape), device=dev)
        bg = mesh.branch_group
        shard = D.RowShard(world=bg.size, rank=bg.rank, group=bg.group)
        gather = lambda: shard.gather(local, 2, tuple(local.shape[1:]), torch.float32, dev)
        gather()
        tg, _ = timed(gather, 20)
        rows = args.sampling_steps * wl.nfe
        res["branch_split"] = dict(
            samples=mesh.dp, ms_single_gpu=t1, ms_pair=t2, speedup_vs_n1=t1 / t2,
            nfe_per_sec=mesh.dp * rows / (t2 * 1e-3), nfe_per_sec_single_gpu=rows / (t1 * 1e-3),
            allgather_us_per_step=tg * 1e3, allgather_bytes=int(local.numel() * 4 * 2),
            max_abs_vs_single_gpu=max_over_ranks((o1 - o2).abs().max().item()),
            equal_on_device=bool(max_over_ranks(0.0 if torch.equal(o1, o2) else 1.0) == 0.0))

    # ---- (ii) the single-sample 200-frame rollout, sharded over chunks x branches
    a = type("A", (), dict(sampling_steps=args.strong_sampling_steps, no_mlp=False, batch=1))()
    wl_long = Workload("re10k_long", a)
    torch.manual_seed(0)
    algo_long = DFoTVideoPose(wl_long.cfg)
    algo_long.diffusion_model.model = algo.diffusion_model.model        # same backbone (weights, packed operands, graphs)
    algo_long = algo_long.to(dev).eval()
    xs_h, conds_h = wl_long.inputs(0)
    xs_l = algo_long._normalize_x(xs_h.to(dev))
    c_l = conds_h.to(dev)
    single = lambda: with_mesh(algo_long, None, 999, lambda: algo_long._predict_videos(xs_l, wl_long.ctx_tokens, c_l))
    shard = lambda: with_mesh(algo_long, mesh, 999, lambda: algo_long.sample_sharded(xs_l, c_l, wl_long.ctx_tokens))
    single()
    shard()
    p0 = algo_long.nfe_rows_planned
    t1, o1 = timed(single, 1)
    rows = algo_long.nfe_rows_planned - p0
    t2, o2 = timed(shard, 1)
    res["long_rollout"] = dict(
        frames=wl_long.gen_frames, sampling_steps=args.strong_sampling_steps, forward_rows=rows, ms_single_gpu=t1,
        ms_sharded=t2, speedup_vs_n1=t1 / t2, nfe_per_sec=rows / (t2 * 1e-3), nfe_per_sec_single_gpu=rows / (t1 * 1e-3),
        generated_frames_per_sec_at_50_steps=wl_long.gen_frames / (t2 * 1e-3 * 50 / args.strong_sampling_steps),
        max_abs_vs_single_gpu=max_over_ranks((o1 - o2).abs().max().item()),
        equal_on_device=bool(max_over_ranks(0.0 if torch.equal(o1, o2) else 1.0) == 0.0))
    # the headline strong-scaling numbers (the long rollout) at the top level, as VERDICT r1 asked
    res.update(nfe_per_sec=res["long_rollout"]["nfe_per_sec"], speedup_vs_n1=res["long_rollout"]["speedup_vs_n1"],
               allgather_us_per_step=res.get("branch_split", {}).get("allgather_us_per_step"))
    return res


# --------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dfot_b200", choices=["dfot_b200", "reference"])
    ap.add_argument("--no-mlp", action="store_true", help="fork default: spatial_mlp_ratio unset (no MLP blocks)")
    ap.add_argument("--workload", default="re10k", choices=["re10k", "k600", "re10k_long", "dmlab"])
    ap.add_argument("--frames", type=int, default=None, help="dmlab: context window length T (16 / 36 / 72 / 144)")
    ap.add_argument("--guidance", type=float, default=None, help="dmlab: vanilla history-guidance scale (> 1: nfe = 2)")
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (default: 4 for re10k, 8 for k600)")
    ap.add_argument("--sampling-steps", type=int, default=50)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-parity", action="store_true", help="skip the one-step GPU-vs-oracle parity check (`parity` key)")
    ap.add_argument("--skip-strong", action="store_true",
                    help="N >= 2: skip the communicating paths (branch split, sharded 200-frame rollout; `strong` key)")
    ap.add_argument("--strong-sampling-steps", type=int, default=4,
                    help="DDIM steps of the 200-frame rollout inside the `strong` section (NFE/s is step-count independent)")
    ap.add_argument("--attn-running-max", action="store_true",
                    help="force the running-maximum attention path (what trained q/k-norm weights with a score bound > 96 "
                         "would select) instead of the bounded-score path the random-init weights allow")
    ap.add_argument("--latency-mode", action="store_true",
                    help="small-batch DiT sampling: split-K block loop, block-per-row AdaLN, single-tile attention items "
                         "(ops.set_latency_mode: a row's bits then depend on its batch; off by default)")
    ap.add_argument("--decode", action="store_true",
                    help="k600 / dmlab: decode the sampled latents to frames inside the e2e region (random-init VideoVAE: "
                         "hidden 128, z 16 / DC-AE: dc_ae_preprocessor.yaml) and report the decode on its own; `value` stays "
                         "sampling-only (SURVEY 8d)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = Workload(args.workload, args)
    cfg = wl.cfg
    strong = wl.name == "re10k_long"          # fixed total work: one rollout spread over the mesh
    br = 2 if (strong and world % 2 == 0) else 1
    config = dict(workload=wl.text, global_batch=wl.batch * (1 if strong else world),
                  parallelism=(f"chunk batches x{world // br} x branches x{br}" if strong
                               else f"samples sharded x{world}"), l2=wl.l2,
                  cuda_graph=True,     # this repo's arm replays the backbone forward from a CUDA graph (both arms name it)
                  attention_path="running-max (forced)" if args.attn_running_max else "by score bound")
    if args.latency_mode:
        config["latency_mode"] = True

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_baseline(wl, args.sampling_steps)
        line = dict(metric="generated_frames_per_sec", value=cb["value"], unit="generated_frames/s", n_gpus=args.gpus,
                    steps=args.steps, warmup=args.warmup, ms_per_step=None, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic", impl="reference", config=config,
                    nfe_per_sec=cb["nfe_per_sec"], cpu_baseline=cb,
                    e2e=dict(value=cb["value"], unit="generated_frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from dfot_b200 import _abi, ops
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.set_latency_mode(args.latency_mode)

    attn_paths = {}
    if True:      # record (and optionally force) the attention kernel variant: bounded scores vs running maximum
        real_attention = ops.attention

        def attention_tagged(qkv, out, R, Ntok, heads, head_dim, score_bound=0.0):
            if args.attn_running_max:
                score_bound = 0.0
            path = "bounded (no running max)" if (0.0 < score_bound <= 96.0 and Ntok > 128) else "running-max"
            attn_paths[f"d{head_dim}_N{Ntok}"] = path
            return real_attention(qkv, out, R, Ntok, heads, head_dim, score_bound=score_bound)

        ops.attention = attention_tagged
    algo = make_weights(cfg, 0).to(dev).eval()
    B = wl.batch
    if args.decode:
        assert wl.name in ("k600", "dmlab"), "--decode applies to the latent workloads (k600: VideoVAE, dmlab: DC-AE)"
        torch.manual_seed(1)
        if wl.name == "k600":
            from dfot_b200.algorithms.vae import VideoVAE
            algo.vae = VideoVAE(hidden_size=128, z_channels=16, embed_dim=16, hidden_size_mult=(1, 2, 4, 4)).to(dev)
        else:
            from dfot_b200.algorithms.vae import MyAutoencoderDC
            algo.vae = MyAutoencoderDC(DCAE_DMLAB_CFG).to(dev)
    xs_host, conds_host = wl.inputs(rank)
    if strong and world > 1:
        from dfot_b200 import distributed as D
        algo.mesh = D.build_mesh(br=br)
    if wl.name.startswith("re10k"):
        xs_host = algo._normalize_x(xs_host.to(dev)).cpu()      # dataset-normalised pixels, as on_after_batch_transfer
    xs_host = xs_host.pin_memory()
    conds_host = None if conds_host is None else conds_host.pin_memory()
    xs_dev = xs_host.to(dev)
    conds_dev = None if conds_host is None else conds_host.to(dev)
    torch.manual_seed(123 + (0 if strong else rank))     # a shared rollout needs one noise stream on every rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_resident():
        if strong:
            return algo.sample_sharded(xs_dev, conds_dev, wl.ctx_tokens)
        return algo._predict_videos(xs_dev, wl.ctx_tokens, conds_dev)

    gathered = []

    def run_e2e():
        # public API with HOST buffers: H2D of the latents, sampling, (N>1: final sample gather), D2H of the result
        batch = {"xs": xs_host.to(dev, non_blocking=True), "gt_videos": None,
                 "conditions": None if conds_host is None else conds_host.to(dev, non_blocking=True)}
        if strong:      # the sharded entry point (every rank ends with the whole video; rank 0's copy goes to the host)
            vids = algo._unnormalize_x(algo.sample_sharded(batch["xs"], batch["conditions"], wl.ctx_tokens))
            return vids.to("cpu")
        vids = algo._sample_all_videos(batch, 0)["prediction"]
        if world > 1:
            if not gathered:
                gathered.extend(torch.empty_like(vids) for _ in range(world))
            dist.all_gather(gathered, vids.contiguous())
        return vids.to("cpu")

    # small-model workload: its per-step working set can sit in the 126 MB L2, so 256 MiB are overwritten between passes
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev) if wl.name == "dmlab" else None

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            if flush is not None:
                flush.zero_()
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    # W >= 3 warm-up steps; a step of the long rollout is 700 sequential backbone calls, one warm-up pass suffices there
    n_warm = max(args.warmup, 1 if strong else 3)
    for _ in range(n_warm):
        run_resident()
    n0 = ops.total_launches()
    p0 = algo.nfe_rows_planned
    with ClockSampler(local_rank) as clocks:
        ms = timed(run_resident, args.steps)
    launches = ops.total_launches() - n0
    rows_planned, passes_counted = algo.nfe_rows_planned - p0, args.steps
    vids_host = run_e2e()
    ms_e2e = timed(run_e2e, args.steps)
    strong_res = None
    if world > 1 and wl.name == "re10k" and not args.skip_strong:
        strong_res = strong_section(args, wl, algo, dev, world)
    ms_decode = None
    if args.decode:       # the decode on its own: latents resident in HBM -> frames in HBM
        lat = algo._unnormalize_x(xs_dev)
        algo._decode(lat)
        ms_decode = timed(lambda: algo._decode(lat), args.steps) / args.steps

    # roofline of the dominant kernel (the tcgen05 GEMM kernel, which also runs the implicit-GEMM convolutions) and
    # of the attention kernel: instrumented extra pass over the same region with CUDA events around every launch
    roof = roof_attn = None
    if rank == 0 or strong:            # (a shared rollout has collectives: every rank takes part in the extra pass)
        ops_gemm, ops_conv, ops_attn = ops.gemm_bf16, ops.conv3x3_bf16, ops.attention
        recs, recs_attn = [], []

        def timed_gemm(a, w, out, epilogue, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            M = kw.get("M") or a.shape[0]
            e0.record()
            ops_gemm(a, w, out, epilogue, **kw)
            e1.record()
            recs.append((2.0 * M * w.shape[0] * w.shape[1], e0, e1))

        def timed_conv(x, w, out, epilogue, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops_conv(x, w, out, epilogue, **kw)
            e1.record()
            recs.append((2.0 * x.shape[0] * x.shape[1] * x.shape[2] * w.shape[0] * 9 * w.shape[3], e0, e1))

        def timed_attn(qkv, out, R, Ntok, heads, head_dim, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops_attn(qkv, out, R, Ntok, heads, head_dim, **kw)
            e1.record()
            recs_attn.append((4.0 * R * heads * Ntok * Ntok * head_dim, e0, e1, f"d{head_dim}_N{Ntok}"))

        ops.gemm_bf16, ops.conv3x3_bf16, ops.attention = timed_gemm, timed_conv, timed_attn
        backbone = algo.diffusion_model.model
        graphs_on = backbone.use_cuda_graph
        backbone.use_cuda_graph = False          # per-launch events need eager launches
        try:
            run_resident()
            torch.cuda.synchronize()
        finally:
            ops.gemm_bf16, ops.conv3x3_bf16, ops.attention = ops_gemm, ops_conv, ops_attn
            backbone.use_cuda_graph = graphs_on
        big = [(f, a.elapsed_time(b)) for f, a, b in recs if f > 1e9]
        flops, dur = sum(f for f, _ in big), sum(d for _, d in big)
        att = [(f, a.elapsed_time(b)) for f, a, b, _ in recs_attn]
        att_by_shape = {}
        for f, a, b, tag in recs_attn:
            d = att_by_shape.setdefault(tag, [0.0, 0.0, 0])
            d[0] += f
            d[1] += a.elapsed_time(b)
            d[2] += 1
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        traffic = {}                     # ncu-measured DRAM bytes of representative launches (committed capture summaries)
        try:
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                traffic = json.load(f)["launches"]
        except Exception:
            pass
        t_gemm = traffic.get("gemm_level3_qkv_16384x3456x1152_bf16" if wl.name.startswith("re10k") else "", {}).get("dram_bytes")
        t_attn = traffic.get("attention_level2_d64_N8192_R8" if wl.name.startswith("re10k") else "", {}).get("dram_bytes")
        ach = flops / (dur * 1e-3) / 1e12
        roof = dict(bound="tensor", kernel="gemm2_bf16_tcgen05_kernel (CTA pair) + gemm_bf16_tcgen05_kernel", achieved=ach, peak=peak, unit="TFLOP/s",
                    frac=ach / peak, traffic=t_gemm, launches=len(big), avg_launch_us=1e3 * dur / max(len(big), 1),
                    peak_source="MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained",
                    gemm_share_of_step=dur / (ms / args.steps),
                    traffic_note="ncu DRAM bytes of the level-3 fused-qkv launch (algorithmic 159 MB); `achieved` averages "
                                 "all GEMM / conv launches, per-shape traffic in profiles/r01_traffic.json")
        if att:
            af, ad = sum(f for f, _ in att), sum(d for _, d in att)
            roof_attn = dict(bound="tensor", kernel="attention_tcgen05_kernel", achieved=af / (ad * 1e-3) / 1e12, peak=peak,
                             unit="TFLOP/s", frac=af / (ad * 1e-3) / 1e12 / peak, traffic=t_attn, launches=len(att),
                             avg_launch_us=1e3 * ad / len(att), share_of_step=ad / (ms / args.steps),
                             by_shape={tag: dict(path=attn_paths.get(tag), tflops=f / (d * 1e-3) / 1e12, launches=n,
                                                 frac=f / (d * 1e-3) / 1e12 / peak)
                                       for tag, (f, d, n) in att_by_shape.items()})
    if world > 1:
        dist.barrier()

    if rank == 0:
        per_step_s = ms / args.steps / 1e3
        frames = B * (1 if strong else world) * wl.gen_frames
        rows = B * world * args.sampling_steps * wl.nfe
        if strong:      # useful forward-rows of one rollout = what a single GPU executes (replicated keyframe rows and
            rows = rows_planned // passes_counted          # noise-only replays are not counted twice)
        per_e2e_s = ms_e2e / args.steps / 1e3
        bytes_out = vids_host.numel() * 4
        bytes_in = xs_host.numel() * 4 + (0 if conds_host is None else conds_host.numel() * 4)
        line = dict(metric="generated_frames_per_sec", value=frames / per_step_s, unit="generated_frames/s",
                    n_gpus=world, steps=args.steps, warmup=n_warm, ms_per_step=ms / args.steps,
                    higher_is_better=True, scaling="strong" if strong else "weak", vs_baseline=None, dtype="bf16",
                    data="synthetic", config=config, nfe_per_sec=rows / per_step_s,
                    model_tflops=rows * wl.forward_row_gflop() / per_step_s / 1e3,
                    e2e=dict(value=frames / per_e2e_s, unit="generated_frames/s", h2d_bytes_per_step=bytes_in,
                             d2h_bytes_per_step=bytes_out, nfe_per_sec=rows / per_e2e_s),
                    gpu_launches=int(launches), clocks=clocks.summary(), roofline=roof)
        if roof_attn is not None:
            line["roofline_attention"] = roof_attn
        if strong_res is not None:
            line["strong"] = strong_res
        if ms_decode is not None:
            n_fr = B * world * vids_host.shape[1]
            line["vae_decode"] = dict(ms_per_batch=ms_decode, decoded_frames_per_sec=n_fr / ms_decode * 1e3,
                                      video_shape=list(vids_host.shape), vae_batch_size=cfg["vae"]["batch_size"],
                                      note=("VideoVAE decoder (hidden 128, mult 1-2-4-4, z 16, random init)" if wl.name == "k600"
                                            else "DC-AE decoder (dc_ae_preprocessor.yaml: 128/256/512/512 channels, 17 blocks, "
                                                 "random init)") + "; e2e includes it, `value` does not (SURVEY 8d excludes "
                                           "the decode)")
        assert bool(algo.diffusion_model.model.use_cuda_graph) == config["cuda_graph"]
        if not args.skip_parity and world == 1 and wl.name != "re10k_long":
            line["parity"] = parity_check(wl, dev)
        if not args.skip_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl, args.sampling_steps)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
