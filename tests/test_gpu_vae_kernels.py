"""GPU: building blocks of the VAE-decode row against plain PyTorch fp32 references of the same
ops — strided GroupNorm over the valid frames of padded clips, the two upsamplers of the causal VideoVAE decoder, pad
filling and the row softmax.  (The causal 3-D convolution is covered in test_gpu_uvit_kernels.py.)"""
import pytest
import torch
import torch.nn.functional as F

from dfot_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
PAD = 2


def rel_err(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def clip_f32(B, T, H, W, C, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((B, PAD + T, H, W, C), generator=g) * 1.5 + 0.3
    return x.to(DEV)


@pytest.mark.parametrize("B,T,H,W,C,silu", [(2, 3, 4, 4, 64, True), (1, 5, 8, 8, 128, False), (3, 9, 16, 16, 32, True),
                                            (2, 17, 32, 32, 64, True)])
def test_groupnorm_over_valid_frames_of_padded_clips(B, T, H, W, C, silu):
    x = clip_f32(B, T, H, W, C, B + T + C)
    g = torch.Generator().manual_seed(7)
    gamma, beta = (1 + 0.1 * torch.randn((C,), generator=g)).to(DEV), (0.1 * torch.randn((C,), generator=g)).to(DEV)
    frame = H * W * C
    sums = torch.empty((B, 32, 3), dtype=torch.float64, device=DEV)
    valid = x.view(-1)[PAD * frame:]                    # clip b's valid frames start at b*(PAD+T)*frame + PAD*frame
    ops.groupnorm_stats_strided(valid, sums, B, T * H * W, (PAD + T) * frame, C)
    out = torch.zeros((B, PAD + T, H, W, C), dtype=torch.bfloat16, device=DEV)
    ops.groupnorm_apply_bf16(valid, sums, gamma, beta, out.view(-1)[PAD * frame:], B, T * H * W, (PAD + T) * frame, C,
                             silu=silu)
    ref = F.group_norm(x[:, PAD:].permute(0, 4, 1, 2, 3), 32, gamma, beta, eps=1e-6)
    ref = (ref * torch.sigmoid(ref) if silu else ref).permute(0, 2, 3, 4, 1)
    assert rel_err(out[:, PAD:], ref) < 4e-3
    assert (out[:, :PAD] == 0).all()                    # pad slots are not touched
    ops.vae_fill_pad_frames(out, B, T, frame)
    assert torch.equal(out[:, 0], out[:, PAD]) and torch.equal(out[:, 1], out[:, PAD])


@pytest.mark.parametrize("B,T,H,W,C", [(2, 3, 4, 4, 64), (1, 1, 8, 8, 32), (1, 5, 16, 16, 128), (2, 2, 8, 4, 64)])
def test_upsamplers_match_torch_interpolate(B, T, H, W, C):
    x = clip_f32(B, T, H, W, C, 3 * B + T)
    v = x[:, PAD:].permute(0, 4, 1, 2, 3)                                  # b c t h w
    # nearest x2 (SpatialUpsample2x)
    out = torch.empty((B, PAD + T, 2 * H, 2 * W, C), dtype=torch.bfloat16, device=DEV)
    ops.vae_upsample2x_bf16(x, out, B, T, H, W, C, temporal=False)
    ref = F.interpolate(v.reshape(B, C * T, H, W), scale_factor=(2, 2), mode="nearest").reshape(B, C, T, 2 * H, 2 * W)
    assert rel_err(out[:, PAD:], ref.permute(0, 2, 3, 4, 1)) < 3e-3
    assert torch.equal(out[:, 0], out[:, PAD]) and torch.equal(out[:, 1], out[:, PAD])
    # first frame bilinear, the rest trilinear (Spatial2xTime2x3DUpsample, causal)
    To = 2 * T - 1
    out = torch.empty((B, PAD + To, 2 * H, 2 * W, C), dtype=torch.bfloat16, device=DEV)
    ops.vae_upsample2x_bf16(x, out, B, T, H, W, C, temporal=True)
    first = F.interpolate(v[:, :, :1], scale_factor=(1, 2, 2), mode="trilinear")
    ref = first if T == 1 else torch.cat([first, F.interpolate(v[:, :, 1:], scale_factor=(2, 2, 2), mode="trilinear")], 2)
    assert ref.shape[2] == To
    assert rel_err(out[:, PAD:], ref.permute(0, 2, 3, 4, 1)) < 3e-3
    assert torch.equal(out[:, 0], out[:, PAD]) and torch.equal(out[:, 1], out[:, PAD])


@pytest.mark.parametrize("rows,n,ld", [(256, 256, 256), (1000, 64, 64), (77, 1024, 1024), (512, 16, 40)])
def test_softmax_rows(rows, n, ld):
    g = torch.Generator().manual_seed(rows + n)
    buf = (torch.randn((rows, ld), generator=g) * 4).to(DEV)
    s = buf[:, :n]
    p = torch.empty((rows, n), dtype=torch.bfloat16, device=DEV)
    ops.softmax_rows_bf16(s, p, scale=0.37)
    ref = torch.softmax(s * 0.37, -1)
    assert (p.float() - ref).abs().max().item() < 4e-3 and rel_err(p, ref) < 4e-3
