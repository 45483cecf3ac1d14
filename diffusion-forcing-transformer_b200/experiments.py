"""Sampling driver — the `validation` task of the reference's experiment layer reduced to what the hot path needs
(experiments/simple_video_generation.py:324-487 `validation` / `run_validation`): build the algorithm from the
`algorithm` config tree, ingest a checkpoint, walk the validation batches through
`on_after_batch_transfer → _sample_all_videos` and collect the videos.  One process per GPU: every batch is sharded
by samples over the `dp x br` mesh of `dfot_b200.distributed` (the reference replicates the model with DDP and lets
Accelerate split the loader; here the split is explicit and the only collective is the final gather of each batch).
Metrics, logging and data modules are outside the scope of this package: batches are plain dicts.

CLI (single GPU, or under torchrun for several) — either the reference's own command line, composed against the user's
checkout of its `configurations/` tree (dfot_b200/hydra_compose.py; `load=<path>` names the checkpoint):
    python -m dfot_b200.experiments --config-dir /path/to/diffusion-forcing-transformer/configurations \
        --input batch.npz --output videos.npz -- dataset=realestate10k_mini algorithm=dfot_video_pose \
        experiment=video_generation @diffusion/continuous load=DFoT_RE10K.ckpt 'experiment.tasks=[validation]' \
        dataset.context_length=1 dataset.n_frames=8 algorithm.tasks.prediction.history_guidance.name=vanilla \
        +algorithm.tasks.prediction.history_guidance.guidance_scale=4.0
or an already resolved `algorithm` tree:
    python -m dfot_b200.experiments --config algo.json --ckpt model.ckpt --input batch.npz --output videos.npz [--br 2]
`batch.npz` holds `videos` (or `latents`) [B, T, C, H, W] in [0, 1] and optionally `conds` [B, T, d].
"""
import argparse
import json
import os
import time
from typing import Dict, Iterable, List, Optional

import torch

from dfot_b200.config import to_config


def build_algo(algorithm_cfg):
    """`compatible_algorithms` of the reference's experiment (simple_video_generation.py:55-59)."""
    from dfot_b200.algorithms.dfot import DFoTVideo, DFoTVideoPose
    cfg = to_config(algorithm_cfg)
    return (DFoTVideoPose if "camera_pose_conditioning" in cfg else DFoTVideo)(cfg)


class SamplingExperiment:
    def __init__(self, algorithm_cfg, ckpt_path: Optional[str] = None, device: Optional[torch.device] = None,
                 branch_group_size: int = 1, manual_seed: Optional[int] = None):
        import torch.distributed as dist
        self.algo = build_algo(algorithm_cfg).eval()
        if ckpt_path:
            self.algo.load_checkpoint(ckpt_path)
        if device is not None:
            self.algo = self.algo.to(device)
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.distributed:
            from dfot_b200 import distributed as D
            self.algo.mesh = D.build_mesh(br=branch_group_size)
        self.rank = dist.get_rank() if self.distributed else 0
        if manual_seed is not None:
            # the reference calls set_seed(manual_seed, device_specific=True) (:341-342) because every rank samples its own
            # batch; here a batch may be REPLICATED over ranks (branch groups; keyframe windows when batch < dp shards), and
            # those ranks must draw the same noise: all ranks share the seed, `sample_sharded` derives a per-shard stream
            # from it only where the samples themselves are sharded.
            torch.manual_seed(manual_seed)
        self.stats = {"batches": 0, "videos": 0, "forward_rows": 0, "seconds": 0.0}

    @torch.no_grad()
    def run_validation(self, batches: Iterable[Dict[str, torch.Tensor]], limit_batch: Optional[int] = None
                       ) -> List[Dict[str, torch.Tensor]]:
        """Returns, per batch, {"gt", "prediction"[, "interpolation"]} un-normalised videos (every rank holds the full
        batch after the final gather, like `accelerator.gather_for_metrics`)."""
        algo, out = self.algo, []
        dev = algo.device
        for i, batch in enumerate(batches):
            if limit_batch is not None and i >= limit_batch:
                break
            batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
            batch = algo.on_after_batch_transfer(batch, i)
            rows0, t0 = algo.nfe_rows, time.perf_counter()
            videos = {"gt": algo._unnormalize_x(batch["xs"]).detach()}
            for task in algo.tasks:
                if task == "prediction":
                    pred = algo.sample_sharded(batch["xs"], batch["conditions"], algo.n_context_tokens)
                else:
                    pred = algo._interpolate_videos(batch["xs"], conditions=batch["conditions"])
                videos[task] = algo._unnormalize_x(pred).detach()
            videos = self._decode_latents(videos, batch)
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            self.stats["seconds"] += time.perf_counter() - t0
            self.stats["forward_rows"] += algo.nfe_rows - rows0
            self.stats["batches"] += 1
            self.stats["videos"] += batch["xs"].shape[0]
            out.append(videos)
        return out

    def _decode_latents(self, videos: Dict[str, torch.Tensor], batch: Dict) -> Dict[str, torch.Tensor]:
        """The latent -> pixel step of `_sample_all_videos` (dfot_video.py:104-111) for latent configurations with a VAE
        configured (`vae.pretrained_path`, or an `algo.vae` attached): every entry is decoded, sharded over all ranks;
        `gt` comes from the dataset's videos when the batch carries them."""
        algo = self.algo
        if not (algo.is_latent_diffusion and (algo.vae is not None or (algo.cfg.get("vae") or {}).get("pretrained_path"))):
            return videos
        from dfot_b200 import distributed as D
        gt = batch.get("gt_videos")
        return {k: (gt if k == "gt" and gt is not None else D.decode_sharded(algo._decode, v)) for k, v in videos.items()}


def _load_tree(path: str):
    with open(path) as f:
        if path.endswith((".yaml", ".yml")):
            import yaml
            return yaml.safe_load(f)
        return json.load(f)


def resolve_cli_config(args):
    """(algorithm tree, checkpoint path) from either `--config` or `--config-dir -- <reference command line>`."""
    if (args.config is None) == (args.config_dir is None):
        raise SystemExit("give exactly one of --config (a resolved algorithm tree) and --config-dir (the reference's tree)")
    if args.config is not None:
        return _load_tree(args.config), args.ckpt
    from dfot_b200.hydra_compose import compose
    cfg = compose(args.config_dir, list(args.overrides))
    tasks = cfg.get("experiment", {}).get("tasks", [])
    if tasks and "validation" not in tasks and "test" not in tasks:
        raise SystemExit(f"experiment.tasks={tasks}: only the sampling tasks (validation / test) are implemented here")
    load = args.ckpt or cfg.get("load")
    if isinstance(load, str) and load.startswith("pretrained:"):
        raise SystemExit(f"load={load}: the reference downloads this from the Hugging Face hub; there is no network here — "
                         "download it yourself and pass the file (load=/path/to/file.ckpt or --ckpt)")
    return cfg["algorithm"], load


def main(argv=None):
    import numpy as np
    import torch.distributed as dist
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", default=None, help="resolved `algorithm` config tree (json / yaml)")
    ap.add_argument("--config-dir", default=None,
                    help="the reference's `configurations/` directory: the arguments after `--` are the reference's own "
                         "command line (group choices, overrides, @shortcuts) and are composed against it")
    ap.add_argument("overrides", nargs="*", help="with --config-dir: the reference's `python -m main` arguments")
    ap.add_argument("--ckpt", default=None, help="reference checkpoint (.ckpt / .safetensors); default: `load=` of the command line")
    ap.add_argument("--input", required=True, help=".npz with videos|latents [B,T,C,H,W] and optional conds [B,T,d]")
    ap.add_argument("--output", required=True)
    ap.add_argument("--br", type=int, default=1, help="branch-group size (splits history-guidance branches over GPUs)")
    ap.add_argument("--seed", type=int, default=None)
    args = ap.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    exp = SamplingExperiment(*resolve_cli_config(args), dev, args.br, args.seed)
    data = np.load(args.input)
    batch = {k: torch.from_numpy(data[k]) for k in data.files}
    videos = exp.run_validation([batch])[0]
    if exp.rank == 0:
        np.savez_compressed(args.output, **{k: v.float().cpu().numpy() for k, v in videos.items()})
        s = exp.stats
        print(json.dumps({"videos": s["videos"], "forward_rows": s["forward_rows"], "seconds": round(s["seconds"], 3),
                          "nfe_per_sec": round(s["forward_rows"] / max(s["seconds"], 1e-9), 2)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
