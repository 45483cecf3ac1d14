"""No-GPU checks of bench.py's workload definitions: the 200-frame RE10K rollout (BASELINE config[3]) must plan exactly
what executing the reference gave in SURVEY.md §8 (96 forward-rows per DDIM step, 14 sequential windows = 700 backbone
calls at 50 steps), and every workload's config must resolve."""
import json
import sys

import pytest
import torch

from helpers import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402
import k4_emulation  # noqa: E402


def _args(**kw):
    base = dict(sampling_steps=2, no_mlp=False, batch=None, frames=None, guidance=None)
    base.update(kw)
    return type("A", (), base)()


def test_long_rollout_rows_and_windows(monkeypatch):
    from dfot_b200 import ops
    from dfot_b200.algorithms.dfot import DFoTVideoPose
    wl = bench.Workload("re10k_long", _args())
    cfg = json.loads(json.dumps(wl.cfg))
    # same plan, toy tensors: the planner only sees frame counts, masks and the task configuration
    cfg["x_shape"] = [3, 16, 16]
    cfg["backbone"].update(channels=[32, 32, 64, 128], emb_channels=64, num_heads=1, num_updown_blocks=[1, 1, 1],
                           num_mid_blocks=1)
    algo = DFoTVideoPose(cfg)
    algo.model_in_dtype = torch.float32
    calls = []

    class Stub(torch.nn.Module):
        def forward(self, x, k, c=None, cm=None, out_dtype=None):
            calls.append(x.shape[0])
            return torch.zeros_like(x, dtype=torch.float32)

    algo.diffusion_model.model = Stub()
    monkeypatch.setattr(ops, "sampler_step_hg", k4_emulation.emulate)
    xs = torch.rand((1, 200, 3, 16, 16))
    out = algo._predict_videos(xs, wl.ctx_tokens, bench.synthetic_poses(1, 200))
    steps = cfg["diffusion"]["sampling_timesteps"]
    assert out.shape == xs.shape
    assert algo.nfe_rows == 96 * steps                       # 2 + 2 keyframe-window rows, 22 + 70 chunk rows per step
    assert len(calls) == 14 * steps                          # 2 windows + 3 + 9 chunk batches, sequential
    assert sorted(set(calls)) == [2, 6, 8]                   # 1 sample x 2 branches; batches of 4 (and one of 3) chunks


@pytest.mark.parametrize("name,kw,nfe,rows", [("re10k", {}, 2, 8), ("k600", {}, 1, 8), ("dmlab", dict(frames=36), 1, 16),
                                              ("dmlab", dict(frames=72, guidance=2.0, batch=4), 2, 8)])
def test_workload_configs_resolve(name, kw, nfe, rows):
    from dfot_b200.config import to_config
    from dfot_b200.algorithms.dfot.history_guidance import HistoryGuidance
    wl = bench.Workload(name, _args(sampling_steps=50, **kw))
    cfg = to_config(wl.cfg)
    hg = HistoryGuidance.from_config(cfg.tasks.prediction.history_guidance, timesteps=cfg.diffusion.timesteps)
    assert wl.nfe == nfe and wl.batch * wl.nfe == rows
    assert hg is not None and wl.forward_row_gflop() > 0
    xs, conds = wl.inputs(0)
    assert xs.shape[:2] == (wl.batch, wl.n_tokens) and tuple(xs.shape[2:]) == tuple(wl.x_shape)
    assert (conds is None) == (name == "k600")
