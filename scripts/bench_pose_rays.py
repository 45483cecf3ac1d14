#!/usr/bin/env python
"""dfot_pose_ray_patches at the RE10K window shape (4 samples x 8 frames, 256x256, patch 2, 15 octaves): 377 MB of bf16."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfot_b200 import ops  # noqa: E402
from dfot_b200.algorithms.dfot.dfot_video_pose import ray_freq_scale  # noqa: E402

frames, res, p = 32, 256, 2
g = res // p
torch.manual_seed(0)
cams = torch.randn((frames, 16), device="cuda")
cams[:, :2] = 200.0 + cams[:, :2].abs()
cams[:, 2:4] = 128.0
fs = ray_freq_scale(15).to("cuda")
out = torch.empty((frames * g * g, p * p * 180), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    ops.pose_ray_patches(cams, fs, out, frames, res, p)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 10
for _ in range(n):
    ops.pose_ray_patches(cams, fs, out, frames, res, p)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / n
print(f"pose_ray_patches {out.numel() * 2 / 1e6:.0f} MB: {us:.1f} us, {out.numel() * 2 / us / 1e6:.2f} TB/s written")
