#!/usr/bin/env python
"""Backbone forward time of the RE10K-size UViT3DPose (CUDA-graph replay) as a function of the number of rows R in one
forward — what the multi-GPU partitioning of the long rollout trades against.
Usage: forward_rows_sweep.py [R ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rows = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 6, 8]
args = type("A", (), dict(sampling_steps=50, no_mlp=False, batch=4))()
wl = bench.Workload("re10k", args)
algo = bench.make_weights(wl.cfg, 0).to("cuda").eval()
model = algo.diffusion_model.model
conds = bench.synthetic_poses(8, 8).to("cuda")
for R in rows:
    cond = algo._window_conditions(conds[:R], 1)
    x = torch.randn((R, 8, 3, 256, 256), device="cuda").to(torch.bfloat16)
    levels = torch.randn((R, 8), device="cuda")
    mask = torch.tensor(([True, False] * R)[:R], device="cuda")
    for _ in range(3):          # eager, capture, replay
        model(x, levels, cond, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 6
    e0.record()
    for _ in range(n):
        model(x, levels, cond, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"R={R}: {ms:.2f} ms per forward, {ms / R:.2f} ms per row, {R / ms * 1e3:.1f} rows/s")
