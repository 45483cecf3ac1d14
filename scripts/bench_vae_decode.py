"""VAE-decode row: time dfot_b200's VideoVAE.decode at the K600 shape — latents [B, 16, 5, 16, 16] ->
[B, 3, 17, 128, 128] (hidden 128, mult (1, 2, 4, 4), random-init weights) — with CUDA events, inputs resident in HBM.
Prints one JSON line; algorithmic FLOPs = 2 * MACs of every convolution / GEMM of the decoder on the valid frames."""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from dfot_b200 import ops  # noqa: E402
from dfot_b200.algorithms.vae import VideoVAE  # noqa: E402


def decoder_flops(m: VideoVAE, B, T, H, W):
    """walks the decoder's topology with the frame counts / resolutions of each level"""
    sd = m.state_dict()
    fl = 0.0

    def conv(name, t, h, w):
        nonlocal fl
        wt = sd[f"{name}.conv.weight"]
        fl += 2.0 * B * t * h * w * wt.numel()

    def res(name, t, h, w):
        conv(f"{name}.conv1", t, h, w); conv(f"{name}.conv2", t, h, w)
        if f"{name}.nin_shortcut.conv.weight" in sd:
            conv(f"{name}.nin_shortcut", t, h, w)

    conv("post_quant_conv", T, H, W); conv("decoder.conv_in", T, H, W)
    res("decoder.mid.block_1", T, H, W); res("decoder.mid.block_2", T, H, W)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", T, H, W)
    C = m.hidden_size * m.mult[-1]
    fl += 2 * 2.0 * B * T * (H * W) ** 2 * C
    for lvl in reversed(range(4)):
        for i in range(m.nrb + 1):
            res(f"decoder.up.{lvl}.block.{i}", T, H, W)
        if lvl >= 1:
            T, H, W = (2 * T - 1 if lvl >= 2 else T), 2 * H, 2 * W
            conv(f"decoder.up.{lvl}.upsample.conv", T, H, W)
    conv("decoder.conv_out", T, H, W)
    return fl, T


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    torch.manual_seed(0)
    m = VideoVAE(hidden_size=128, z_channels=16, embed_dim=16, hidden_size_mult=(1, 2, 4, 4)).cuda()
    z = torch.randn((a.batch, 16, 5, 16, 16), device="cuda")
    fl, frames = decoder_flops(m, a.batch, 5, 16, 16)
    for _ in range(a.warmup):
        v = m.decode(z, 17)
    torch.cuda.synchronize()
    n0 = ops.total_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        v = m.decode(z, 17)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"metric": "VAE-decoded frames/s (VideoVAE decoder, K600 shape)", "value": a.batch * frames / ms * 1e3,
                      "unit": "frames/s", "ms_per_decode": ms, "batch": a.batch, "frames": frames,
                      "video_shape": list(v.shape), "algorithmic_tflop": fl / 1e12,
                      "achieved_tflops": fl / ms / 1e9, "gpu_launches_per_decode": (ops.total_launches() - n0) // a.steps,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "dtype": "bf16 operands, fp32 accumulate/residual"}))


if __name__ == "__main__":
    main()
