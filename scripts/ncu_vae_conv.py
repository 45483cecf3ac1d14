"""One launch shape of the VAE-decode row for ncu: the causal 3x3x3 convolution of the VideoVAE decoder's last level at the
K600 shape — 8 clips x (2 + 17) frames of 128 x 128, 128 -> 128 channels, fp32 residual epilogue (conv2 of a ResnetBlock3D).
Plain run: CUDA-event time of the launch (after warm-up).   ncu: -k regex:gemm -c 1 after the warm-up launches (-s 3)."""
import sys

import torch

sys.path.insert(0, ".")
from dfot_b200 import ops  # noqa: E402

B, T, H, W, C = 8, 17, 128, 128, 128
n_all = B * (2 + T)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((n_all, H, W, C), device="cuda", generator=g).to(torch.bfloat16)
w = (torch.randn((C, 3, 3, 3, C), device="cuda", generator=g) / (27 * C) ** 0.5).to(torch.bfloat16)
bias = torch.zeros((C,), device="cuda")
resid = torch.randn(((n_all - 2) * H * W, C), device="cuda", generator=g)
out = torch.empty_like(resid)
for _ in range(3):
    ops.conv3d_causal_bf16(x, w, out, ops.EPI_RESID_F32, bias=bias, resid=resid)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 1 if len(sys.argv) > 1 else 10
e0.record()
for _ in range(n):
    ops.conv3d_causal_bf16(x, w, out, ops.EPI_RESID_F32, bias=bias, resid=resid)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
M = (n_all - 2) * H * W
flop = 2.0 * M * 27 * C * C
byt = x.numel() * 2 + resid.numel() * 4 + out.numel() * 4
print(f"conv3d_causal 128->128 @128x128 x {n_all - 2} frames: {ms * 1e3:.1f} us, {flop / ms / 1e9:.1f} TFLOP/s, "
      f"algorithmic bytes {byt / 1e6:.0f} MB -> {byt / ms / 1e6:.0f} GB/s")
