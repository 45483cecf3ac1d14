"""TEST INFRASTRUCTURE — the reference's README command lines composed against the reference's own `configurations/`
tree (authoring container only; neither Hydra nor OmegaConf is installed here, so the composer under test,
dfot_b200/hydra_compose.py, does the composing — its semantics are pinned separately on a hand-checked miniature tree in
tests/test_hydra_compose.py):
    python -m oracle.make_goldens_config
Writes tests/golden/composed_configs.json: for every command the resolved `algorithm` tree.  The test then pins the
hand-resolved configurations of bench.py (`re10k_cfg`, `k600_cfg`, `dmlab_cfg`, the 200-frame rollout) to these trees on
every key the sampling path reads."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dfot_b200.hydra_compose import compose  # noqa: E402

CONFIG_DIR = "/root/reference/configurations"
COMMANDS = {      # README.md:69, 74, 98 and BASELINE.json configs[4]
    "re10k_short": "+name=single_image_to_short dataset=realestate10k_mini algorithm=dfot_video_pose experiment=video_generation "
                   "@diffusion/continuous load=pretrained:DFoT_RE10K.ckpt experiment.tasks=[validation] "
                   "experiment.validation.data.shuffle=True dataset.context_length=1 dataset.frame_skip=20 dataset.n_frames=8 "
                   "experiment.validation.batch_size=1 algorithm.tasks.prediction.history_guidance.name=vanilla "
                   "+algorithm.tasks.prediction.history_guidance.guidance_scale=4.0",
    "re10k_long": "+name=single_image_to_long dataset=realestate10k_mini algorithm=dfot_video_pose experiment=video_generation "
                  "@diffusion/continuous load=pretrained:DFoT_RE10K.ckpt experiment.tasks=[validation] "
                  "experiment.validation.data.shuffle=True dataset.context_length=1 dataset.frame_skip=1 dataset.n_frames=200 "
                  "algorithm.tasks.prediction.keyframe_density=0.0625 algorithm.tasks.interpolation.max_batch_size=4 "
                  "experiment.validation.batch_size=1 algorithm.tasks.prediction.history_guidance.name=stabilized_vanilla "
                  "+algorithm.tasks.prediction.history_guidance.guidance_scale=4.0 "
                  "+algorithm.tasks.prediction.history_guidance.stabilization_level=0.02 "
                  "algorithm.tasks.interpolation.history_guidance.name=vanilla "
                  "+algorithm.tasks.interpolation.history_guidance.guidance_scale=1.5",
    "k600": "+name=K600 dataset=kinetics_600 algorithm=dfot_video experiment=video_generation @DiT/XL",
    "dmlab_T36": "+name=DMLAB dataset=dmlab algorithm=dfot_video experiment=video_generation @diffusion/continuous @DiT/B "
                 "dataset.max_frames=36",
}


def main():
    out = {}
    for name, cmd in COMMANDS.items():
        cfg = compose(CONFIG_DIR, cmd.split())
        out[name] = dict(argv=cmd.split(), algorithm=cfg["algorithm"], experiment_tasks=cfg["experiment"]["tasks"],
                         load=cfg.get("load"))
        print(name, cfg["algorithm"]["backbone"]["name"], cfg["algorithm"]["x_shape"], cfg["algorithm"]["n_frames"])
    with open(os.path.join(ROOT, "tests", "golden", "composed_configs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
