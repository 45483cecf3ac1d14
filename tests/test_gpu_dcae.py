"""-m gpu: the DC-AE decoder on the B200 kernels (DMLab / Minecraft latents): every glue kernel of
csrc/dcae.cu against plain torch fp32, and `MyAutoencoderDC.decode` against the fixture of the executed reference and
against the oracle at the full DMLab topology (configurations/algorithm/dc_ae_preprocessor.yaml).
Tolerance (bf16 conv / GEMM operands, fp32 accumulation and residual stream): relative L2 <= 2e-2, PSNR >= 40 dB."""
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from dfot_b200 import ops  # noqa: E402
from dfot_b200.algorithms.vae import MyAutoencoderDC  # noqa: E402
from helpers import GOLDEN  # noqa: E402
from oracle.dc_ae import DCAEDecoderOracle, decoder_param_shapes, seeded_weights, small_cfg  # noqa: E402

DEV = "cuda"


def _psnr(a, b):
    rng = (b.max() - b.min()).item()
    return 10 * math.log10(rng * rng / max(((a - b) ** 2).mean().item(), 1e-30))


def test_relu_and_pixel_shuffle():
    g = torch.Generator().manual_seed(0)
    x = torch.randn((4, 8, 8, 64), generator=g).to(DEV).to(torch.bfloat16)
    y = x.clone()
    ops.relu_bf16(y)
    assert torch.equal(y, F.relu(x.float()).to(torch.bfloat16))
    n, H, W, C, Cx = 3, 8, 8, 16, 32                           # repeats = 4C / Cx = 2; conv rows padded to 72 columns
    conv = torch.randn((n * H * W, 72), generator=g).to(DEV)
    sc = torch.randn((n * H * W, Cx), generator=g).to(DEV)
    of = torch.empty((n, 2 * H, 2 * W, C), device=DEV)
    ob = torch.empty((n, 2 * H, 2 * W, C), device=DEV, dtype=torch.bfloat16)
    ops.pixel_shuffle2x(conv, C, n, H, W, shortcut=sc, repeats=2, out_f32=of, out_bf16=ob)
    a = conv[:, : 4 * C].reshape(n, H, W, 4 * C).permute(0, 3, 1, 2)
    b = sc.reshape(n, H, W, Cx).permute(0, 3, 1, 2).repeat_interleave(2, dim=1)
    ref = (F.pixel_shuffle(a, 2) + F.pixel_shuffle(b, 2)).permute(0, 2, 3, 1)
    assert torch.equal(of, ref) and torch.equal(ob, ref.to(torch.bfloat16))
    ops.pixel_shuffle2x(conv, C, n, H, W, out_f32=of)
    assert torch.equal(of, F.pixel_shuffle(a, 2).permute(0, 2, 3, 1))


@pytest.mark.parametrize("heads,d,HW", [(16, 32, 64), (4, 32, 256), (3, 16, 100)])
def test_linear_attention_relu(heads, d, HW):
    n = 5
    g = torch.Generator().manual_seed(HW)
    qkv = torch.randn((n * HW, 3 * heads * d + 8), generator=g).to(DEV)[:, : 3 * heads * d]      # strided rows
    out = torch.empty((n * HW, heads * d), device=DEV)
    ops.linear_attention_relu(qkv, out, n, HW, heads, d, 1e-15)
    t = qkv.reshape(n, HW, heads, 3 * d).permute(0, 2, 3, 1)                                     # [n, heads, 3d, HW]
    q, k, v = t.chunk(3, dim=2)
    q, k = F.relu(q), F.relu(k)
    v = F.pad(v, (0, 0, 0, 1), value=1)
    h = torch.matmul(torch.matmul(v, k.transpose(-1, -2)), q)
    ref = (h[:, :, :-1] / (h[:, :, -1:] + 1e-15)).reshape(n, heads * d, HW).permute(0, 2, 1).reshape(n * HW, heads * d)
    assert (out - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())


def test_dwconv_glu_and_rmsnorm_affine():
    n, H, W, Ch = 3, 8, 8, 64
    g = torch.Generator().manual_seed(2)
    x = torch.randn((n, H, W, 2 * Ch), generator=g).to(DEV).to(torch.bfloat16)
    w, b = torch.randn((2 * Ch, 9), generator=g).to(DEV) / 3, torch.randn((2 * Ch,), generator=g).to(DEV)
    out = torch.empty((n, H, W, Ch), device=DEV, dtype=torch.bfloat16)
    ops.dwconv3x3_glu_bf16(x, w, b, out, n, H, W, Ch)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.reshape(2 * Ch, 1, 3, 3), b, padding=1, groups=2 * Ch)
    h, gate = y.chunk(2, dim=1)
    ref = (h * F.silu(gate)).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
    M, C = 200, 96
    xx, ww, bb = torch.randn((M, C + 8), generator=g).to(DEV)[:, :C], torch.randn((C,), generator=g).to(DEV), torch.randn((C,), generator=g).to(DEV)
    res = torch.randn((M, C), generator=g).to(DEV)
    of = torch.empty((M, C), device=DEV)
    ob = torch.empty((M, C), device=DEV, dtype=torch.bfloat16)
    ops.rmsnorm_affine(xx, ww, bb, 1e-5, resid=res, out_f32=of)
    ref = xx * torch.rsqrt(xx.pow(2).mean(-1, keepdim=True) + 1e-5) * ww + bb
    assert (of - (ref + res)).abs().max().item() <= 1e-5
    ops.rmsnorm_affine(xx, ww, bb, 1e-5, relu=True, out_bf16=ob)
    assert (ob.float() - F.relu(ref)).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


def test_decode_matches_the_executed_reference_fixture():
    with open(os.path.join(GOLDEN, "dcae_decode.json")) as f:
        meta = json.load(f)
    arr = np.load(os.path.join(GOLDEN, "dcae_decode.npz"))
    sd = seeded_weights(decoder_param_shapes(meta["cfg"]), meta["weight_seed"])
    vae = MyAutoencoderDC(meta["cfg"])
    vae.load_state_dict(sd)
    vae = vae.to(DEV)
    n0 = ops.total_launches()
    out = vae.decode(torch.from_numpy(arr["z"]).to(DEV)).cpu()
    assert ops.total_launches() > n0
    ref = torch.from_numpy(arr["image"])
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel <= 2e-2 and _psnr(out, ref) >= 40.0, (rel, _psnr(out, ref))
    again = vae.decode(torch.from_numpy(arr["z"]).to(DEV)).cpu()
    assert torch.equal(out, again)                                               # deterministic (no float atomics)


def test_decode_at_the_dmlab_topology_vs_oracle():
    """dc_ae_preprocessor.yaml: channels 128/256/512/512, layers 0/5/10/2 (17 blocks), 32 latent channels, 8x8 -> 64x64."""
    cfg = dict(small_cfg(), latent_channels=32, decoder_block_out_channels=[128, 256, 512, 512],
               decoder_layers_per_block=[0, 5, 10, 2])
    sd = seeded_weights(decoder_param_shapes(cfg), 5)
    g = torch.Generator().manual_seed(6)
    z = torch.randn((4, 32, 8, 8), generator=g)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = DCAEDecoderOracle(sd, cfg).decode(z)
    vae = MyAutoencoderDC(cfg)
    vae.load_state_dict(sd)
    out = vae.to(DEV).decode(z.to(DEV)).cpu()
    rel = ((out - ref).norm() / ref.norm()).item()
    print(f"DC-AE DMLab topology: rel L2 {rel:.3e}, PSNR {_psnr(out, ref):.1f} dB")
    assert out.shape == (4, 3, 64, 64) and rel <= 2e-2 and _psnr(out, ref) >= 40.0
