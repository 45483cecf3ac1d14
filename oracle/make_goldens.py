"""TEST INFRASTRUCTURE — generate tests/golden/* by EXECUTING the reference.

Run in the authoring container only (needs /root/reference):
    python -m oracle.make_goldens
The reference ships no golden vectors for the sampling path, so
these fixtures are produced by importing the reference itself through
oracle/ref_shim.py.  Outputs:
  tests/golden/integers.json        scheduling matrices, ddim level tables, HG branch tables,
                                    interpolation plans, mask evolution (bit-exact artefacts)
  tests/golden/schedules.npz        fp32 diffusion buffers (alphas_cumprod, logsnr)
  tests/golden/weights_<w>.npz      random-init reference state dicts (zero params re-drawn)
  tests/golden/case_<name>.npz      inputs, per-step traces and final rollouts of the reference
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.cases import algorithm_cfg, continuous_overrides, golden_cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED, REDRAW_SEED, DATA_SEED, SAMPLING_SEED = 0, 1, 123, 321


def build_reference_algo(cfg: dict):
    from algorithms.dfot.dfot_video import DFoTVideo
    from algorithms.dfot.dfot_video_pose import DFoTVideoPose
    torch.manual_seed(WEIGHT_SEED)
    cls = DFoTVideoPose if cfg["backbone"]["name"] == "u_vit3d_pose" else DFoTVideo
    algo = cls(ref_shim.to_dc(cfg)).eval()
    ref_shim.rerandomize_zero_params(algo, REDRAW_SEED)
    return algo


def synthetic_inputs(cfg: dict, batch: int):
    g = torch.Generator().manual_seed(DATA_SEED)
    xs = torch.randn((batch, cfg["n_frames"], *cfg["x_shape"]), generator=g)
    conds = None
    if "camera_pose_conditioning" in cfg:
        from oracle.cases import synthetic_poses
        conds = synthetic_poses(batch, cfg["n_frames"])
    elif cfg["external_cond_dim"]:
        conds = torch.randn((batch, cfg["n_frames"], cfg["external_cond_dim"]), generator=g)
    return xs, conds


def trace_reference(algo):
    """Record every diffusion_model.sample_step call (inputs, backbone output, result)."""
    dm = algo.diffusion_model
    trace = []
    orig_step, orig_model = dm.sample_step, dm.model.forward
    last = {}

    def model_fwd(x, k, cond=None, cond_mask=None):
        out = orig_model(x, k, cond, cond_mask)
        last.update(model_level=k.clone(), model_out=out.clone())
        return out

    def step(x, curr_noise_level, next_noise_level, external_cond, external_cond_mask=None, guidance_fn=None):
        res = orig_step(x, curr_noise_level, next_noise_level, external_cond, external_cond_mask, guidance_fn)
        trace.append(dict(model_in=x.clone(), levels_from=curr_noise_level.clone(),
                          levels_to=next_noise_level.clone(),
                          cond_mask=None if external_cond_mask is None else external_cond_mask.clone(),
                          model_level=last["model_level"], model_out=last["model_out"], step_out=res.clone()))
        return res

    dm.sample_step = step
    dm.model.forward = model_fwd
    return trace


def run_case(name: str, spec: dict, weights_out: dict):
    cfg = spec["cfg"]
    algo = build_reference_algo(cfg)
    w = spec["weights"]
    sd = {k[len("diffusion_model.model."):]: v.detach().numpy() for k, v in algo.state_dict().items()
          if k.startswith("diffusion_model.model.")}
    if w in weights_out:
        for k, v in sd.items():  # same architecture + seeds must give identical weights
            assert np.array_equal(weights_out[w][k], v), (name, k)
    weights_out[w] = sd
    xs, conds = synthetic_inputs(cfg, spec["batch"])
    trace = trace_reference(algo)
    torch.manual_seed(SAMPLING_SEED)
    with torch.no_grad():
        out = algo._predict_videos(xs.clone(), n_context_tokens=cfg["context_frames"], conditions=conds)
    arrays = {"xs": xs.numpy(), "prediction": out.numpy()}
    if conds is not None:
        arrays["conds"] = conds.numpy()
    for i, t in enumerate(trace):
        for k, v in t.items():
            if v is not None:
                arrays[f"step{i:03d}.{k}"] = v.numpy()
    arrays["n_steps"] = np.array(len(trace))
    np.savez_compressed(os.path.join(OUT, f"case_{name}.npz"), **arrays)
    with open(os.path.join(OUT, f"case_{name}.json"), "w") as f:
        json.dump(dict(cfg=cfg, batch=spec["batch"], weights=w, sampling_seed=SAMPLING_SEED,
                       data_seed=DATA_SEED), f, indent=1)
    print(f"case {name}: {len(trace)} sample_step calls, prediction {tuple(out.shape)}")


def integer_goldens():
    """Bit-exact artefacts, produced by the reference's own functions."""
    from algorithms.dfot.history_guidance import HistoryGuidance
    out = {}
    # --- M1: scheduling matrices + ddim tables
    sched = []
    for kind in ["full_sequence", "autoregressive", "interleaved", "gibbs"]:
        for horizon, padding, steps in [(8, 0, 10), (4, 0, 4), (3, 2, 4), (5, 3, 50), (4, 0, 6), (8, 0, 50)]:
            if kind == "gibbs" and steps > 10:
                continue
            cfg = algorithm_cfg(**{"scheduling_matrix": kind, "diffusion.sampling_timesteps": steps,
                                   "backbone.hidden_size": 64, "backbone.depth": 1, "backbone.num_heads": 1,
                                   "x_shape": [4, 8, 8], "max_frames": horizon + padding})
            algo = build_reference_algo(cfg)
            m = algo._generate_scheduling_matrix(horizon, padding)
            sched.append(dict(kind=kind, horizon=horizon, padding=padding, steps=steps, matrix=m.tolist()))
    out["scheduling_matrices"] = sched
    ddim = []
    for steps in [1, 2, 3, 4, 5, 6, 10, 20, 25, 50, 100, 250, 999]:
        cfg = algorithm_cfg(**{"diffusion.sampling_timesteps": steps, "backbone.hidden_size": 64,
                               "backbone.depth": 1, "backbone.num_heads": 1, "x_shape": [4, 8, 8]})
        algo = build_reference_algo(cfg)
        lv = algo.diffusion_model.ddim_idx_to_noise_level(torch.arange(steps + 1))
        ddim.append(dict(steps=steps, levels=lv.tolist()))
    out["ddim_levels"] = ddim
    # --- H3: branch tables
    schemes = [
        dict(name="conditional"), dict(name="stabilized_conditional", stabilization_level=0.02),
        dict(name="vanilla", guidance_scale=4.0), dict(name="vanilla", guidance_scale=1.0),
        dict(name="vanilla", guidance_scale=2.0, use_external_cond_guidance=False),
        dict(name="stabilized_vanilla", guidance_scale=4.0, stabilization_level=0.02),
        dict(name="stabilized_vanilla", guidance_scale=4.0, stabilization_level=0.1),
        dict(name="fractional", guidance_scale=4.0, freq_scale=0.3),
        dict(name="stabilized_fractional", guidance_scale=4.0, freq_scale=0.3, stabilization_level=0.02),
        dict(name="stabilized_fractional", guidance_scale=1.5, freq_scale=0.07, stabilization_level=0.05),
        dict(name="temporal", hist_subsequences=[[0], [1, 2]], hist_weights=[2, 2]),
        dict(name="temporal", hist_subsequences=[[0, -1], [1]], hist_weights=[1.0, 0.5], gen_segments=[[0, 1], [2]]),
        dict(name="custom", hist_segments=[dict(time_indices="all", freq_ranges=[[0.0, 0.5], [0.2, 1.0]]),
                                           dict(time_indices=[0, 2], freq_ranges=["all"],
                                                freq_ranges_if_generated=[[0.001, 1.0]])],
             hist_weights=[1.0, 2.0]),
    ]
    masks = [[1, 2, 2, 0, 0, 0, -1, -1], [1, 0, 0, 0, 0, 0, 0, 0], [1, 1, 1, 0, 0, 0, 0, 0],
             [2, 2, 2, 2, 0, 0, 0, 0], [1, 0, 0, 1, 0, 0, 1, 1], [1, 2, 2, 2, 2, 2, 2, 2], [0, 0, 0, 0]]
    tables = []
    for sc in schemes:
        for mask in masks:
            hgd = HistoryGuidance.from_config(ref_shim.to_dc(dict(sc, visualize=False)), timesteps=1000)
            m = torch.tensor([mask, mask])
            try:
                with hgd(m) as mgr:
                    rec = dict(scheme=sc, mask=mask, manager=type(mgr).__name__, nfe=mgr.nfe)
                    if type(mgr).__name__ == "HistoryGuidanceManager":
                        rec.update(hist_indices=mgr.hist_indices.tolist(), gen_indices=mgr.gen_indices.tolist(),
                                   gen_mask=mgr.gen_mask.long().tolist(),
                                   hist_noise_levels=mgr.hist_noise_levels.tolist(),
                                   cond_mask=mgr.cond_mask.long().tolist(), weights=mgr.weights.tolist())
            except (AssertionError, IndexError) as e:
                rec = dict(scheme=sc, mask=mask, error=type(e).__name__)
            tables.append(rec)
    out["hg_branch_tables"] = tables
    # --- interpolation plans (dfot_video.py:219-261), extracted by intercepting _sample_sequence
    plans = []
    for n_frames, density, max_frames in [(9, 0.5, 4), (200, 0.0625, 8), (33, 0.25, 8), (17, 0.2, 8), (12, 0.34, 6)]:
        cfg = algorithm_cfg(**{"backbone.hidden_size": 64, "backbone.depth": 1, "backbone.num_heads": 1,
                               "x_shape": [4, 8, 8], "max_frames": max_frames, "n_frames": n_frames,
                               "diffusion.sampling_timesteps": 2})
        algo = build_reference_algo(cfg)
        T = n_frames
        keys = torch.linspace(0, T - 1, round(density * T)).round().long()
        keys = torch.cat([torch.arange(1), keys]).unique()
        known = torch.zeros((1, T), dtype=torch.bool)
        known[:, keys] = True
        calls = []

        def fake(batch_size, context=None, context_mask=None, conditions=None, history_guidance=None, pbar=None,
                 **kw):
            calls.append(dict(batch=batch_size, mask=context_mask.tolist()))
            return context.clone(), None

        algo._sample_sequence = fake
        frames_seen = []
        orig_pad = algo._pad_to_max_tokens

        def pad(y):
            return orig_pad(y)

        algo._pad_to_max_tokens = pad
        with torch.no_grad():
            algo._interpolate_videos(torch.zeros((1, T, 4, 8, 8)), context_mask=known.clone(),
                                     conditions=torch.zeros((1, T, 1)) if False else None)
        plans.append(dict(n_frames=n_frames, density=density, max_tokens=max_frames, keyframes=keys.tolist(),
                          calls=calls))
    out["interpolation_calls"] = plans
    return out


def float_goldens():
    arrays = {}
    for tag, over in [("cosine", {}), ("continuous", continuous_overrides()),
                      ("sigmoid_zt", {"diffusion.beta_schedule": "sigmoid", "diffusion.schedule_fn_kwargs": {}}),
                      ("cosine_shift", {"diffusion.schedule_fn_kwargs": dict(shift=0.5)})]:
        cfg = algorithm_cfg(**{"backbone.hidden_size": 64, "backbone.depth": 1, "backbone.num_heads": 1,
                               "x_shape": [4, 8, 8], **over})
        dm = build_reference_algo(cfg).diffusion_model
        for b in ["alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"]:
            arrays[f"{tag}.{b}"] = getattr(dm, b).numpy()
        if hasattr(dm, "logsnr"):
            arrays[f"{tag}.logsnr"] = dm.logsnr.numpy()
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **arrays)


def main():
    ref_shim.install()
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "integers.json"), "w") as f:
        json.dump(integer_goldens(), f)
    float_goldens()
    weights = {}
    for name, spec in golden_cases().items():
        run_case(name, spec, weights)
    for w, sd in weights.items():
        np.savez_compressed(os.path.join(OUT, f"weights_{w}.npz"), **sd)
    print("weights:", {w: sum(v.size for v in sd.values()) for w, sd in weights.items()})


if __name__ == "__main__":
    main()
