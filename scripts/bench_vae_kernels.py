"""HBM-bound kernels of the VAE-decode row at the K600 decoder's shapes (8 clips): CUDA-event time per launch and
algorithmic GB/s (bytes a perfect implementation moves: each input element read once, each output written once).
The working sets (0.3-2.6 GB) exceed the 126 MB L2, so no flush is needed between launches."""
import json
import sys

import torch

sys.path.insert(0, ".")
from dfot_b200 import ops  # noqa: E402

DEV, PAD = "cuda", 2


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def line(name, ms, byt, **kw):
    print(json.dumps(dict(kernel=name, us=round(ms * 1e3, 1), algorithmic_MB=round(byt / 1e6, 1), GBps=round(byt / ms / 1e6), **kw)))


def main():
    B = 8
    for (T, H, C) in [(17, 128, 128), (17, 64, 256), (9, 32, 512)]:
        frame = H * H * C
        x = torch.randn((B, PAD + T, H, H, C), device=DEV)
        y = torch.empty((B, PAD + T, H, H, C), device=DEV, dtype=torch.bfloat16)
        sums = torch.empty((B, 32, 3), dtype=torch.float64, device=DEV)
        g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        xv, yv = x.view(-1)[PAD * frame:], y.view(-1)[PAD * frame:]
        n_valid = B * T * frame
        ms = timed(lambda: ops.groupnorm_stats_strided(xv, sums, B, T * H * H, (PAD + T) * frame, C))
        line("groupnorm_stats_strided f32", ms, 4 * n_valid, shape=[B, T, H, H, C])
        ms = timed(lambda: ops.groupnorm_apply_bf16(xv, sums, g, b, yv, B, T * H * H, (PAD + T) * frame, C))
        line("groupnorm_apply_bf16 (+SiLU)", ms, 6 * n_valid, shape=[B, T, H, H, C])
        ms = timed(lambda: ops.vae_fill_pad_frames(y, B, T, frame))
        line("vae_fill_pad_frames", ms, 3 * 2 * B * frame, shape=[B, T, H, H, C])
        del x, y
    # upsamplers: level 3 -> 2 (trilinear, 512 ch), 2 -> 1 (trilinear, 512 ch), 1 -> 0 (nearest, 256 ch)
    for (T, H, C, temporal) in [(5, 16, 512, True), (9, 32, 512, True), (17, 64, 256, False)]:
        To = 2 * T - 1 if temporal else T
        x = torch.randn((B, PAD + T, H, H, C), device=DEV)
        out = torch.empty((B, PAD + To, 2 * H, 2 * H, C), device=DEV, dtype=torch.bfloat16)
        ms = timed(lambda: ops.vae_upsample2x_bf16(x, out, B, T, H, H, C, temporal))
        line("vae_upsample2x_bf16 " + ("trilinear" if temporal else "nearest"), ms, 4 * B * T * H * H * C + 2 * out.numel(),
             shape_in=[B, T, H, H, C])
    rows, n = B * 7 * 256, 256
    s = torch.randn((rows, n), device=DEV)
    p = torch.empty((rows, n), device=DEV, dtype=torch.bfloat16)
    ms = timed(lambda: ops.softmax_rows_bf16(s, p, scale=512 ** -0.5))
    line("softmax_rows_bf16", ms, 6 * rows * n, rows=rows, n=n, note="14 MB working set: L2-resident")


if __name__ == "__main__":
    main()
