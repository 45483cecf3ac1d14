#!/usr/bin/env python
"""Eager forwards of a DiT3D workload's backbone (ncu launch-list target for the latency-bound small-batch cases).
Usage: profile_dit_forward.py dmlab|k600 [batch] [frames] [n_forwards]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dmlab"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
n = int(sys.argv[4]) if len(sys.argv) > 4 else 3
args = type("A", (), dict(sampling_steps=50, no_mlp=False, batch=B, frames=T, guidance=None))()
wl = bench.Workload(name, args)
algo = bench.make_weights(wl.cfg, 0).to("cuda").eval()
model = algo.diffusion_model.model
model.use_cuda_graph = "--graph" in sys.argv
xs, conds = wl.inputs(0)
x = xs.to("cuda").to(torch.bfloat16)
levels = torch.randn((B, wl.n_tokens), device="cuda")
cond = None if conds is None else algo._process_conditions(conds.to("cuda"))
for i in range(n + 2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model(x, levels, cond, None)
    e1.record()
    torch.cuda.synchronize()
    print(f"forward {i}: {e0.elapsed_time(e1) * 1e3:.1f} us for {B} rows x {wl.n_tokens} frames")
