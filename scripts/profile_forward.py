#!/usr/bin/env python
"""Two eager forwards of the RE10K-size UViT3DPose backbone (R = batch*2 rows) — the ncu launch-list target.
Usage: profile_forward.py [batch] [n_forwards]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
args = type("A", (), dict(sampling_steps=50, no_mlp=False, batch=B))()
wl = bench.Workload("re10k", args)
algo = bench.make_weights(wl.cfg, 0).to("cuda").eval()
model = algo.diffusion_model.model
model.use_cuda_graph = False
xs, conds = wl.inputs(0)
cond = algo._window_conditions(conds.to("cuda"), 2)
x = torch.randn((2 * B, 8, 3, 256, 256), device="cuda").to(torch.bfloat16)
levels = torch.randn((2 * B, 8), device="cuda")
mask = torch.tensor([True, False] * B, device="cuda")
for i in range(n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model(x, levels, cond, mask)
    e1.record()
    torch.cuda.synchronize()
    print(f"forward {i}: {e0.elapsed_time(e1):.2f} ms for {2 * B} rows")
