"""GPU: the VAE-decode row — dfot_b200's VideoVAE.decode on the B200 kernels against
(i) the fixture produced by executing the reference's VideoVAE (tests/golden/vae_video_decode.npz, made by
oracle/make_goldens_vae.py), (ii) the CPU oracle on a wider seeded case, and (iii) at the full K600 shape, the
oracle's torch fp32 ops executed on the GPU plus batch independence.

Tolerance: bf16 conv operands with fp32 accumulation and an fp32 residual stream against the reference's fp32 —
relative L2 error <= 2e-2 and PSNR >= 40 dB over the decoded range (BASELINE.json's bar for final videos)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from dfot_b200.algorithms.vae import VideoVAE
from helpers import GOLDEN
from oracle.video_vae import VideoVAEDecoderOracle, decoder_param_shapes, seeded_weights

pytestmark = pytest.mark.gpu
DEV = "cuda"
REL_TOL, PSNR_MIN = 2e-2, 40.0


def _errors(got: torch.Tensor, ref: torch.Tensor):
    got, ref = got.float().cpu(), ref.float().cpu()
    rel = ((got - ref).norm() / ref.norm()).item()
    span = (ref.max() - ref.min()).item()
    psnr = 10 * math.log10(span ** 2 / max(((got - ref) ** 2).mean().item(), 1e-30))
    return rel, psnr


def _model(hidden, z_ch, embed, mult, seed):
    shapes = decoder_param_shapes(hidden, z_ch, embed, mult)
    sd = seeded_weights(shapes, seed)
    m = VideoVAE(hidden_size=hidden, z_channels=z_ch, embed_dim=embed, hidden_size_mult=mult)
    m.load_state_dict(sd)
    return m.to(DEV), sd


def test_decode_matches_reference_fixture():
    with open(os.path.join(GOLDEN, "vae_video_decode.json")) as f:
        c = json.load(f)["case"]
    arr = dict(np.load(os.path.join(GOLDEN, "vae_video_decode.npz")))
    m, _ = _model(c["hidden_size"], c["z_channels"], c["embed_dim"], tuple(c["hidden_size_mult"]), c["weight_seed"])
    z = torch.from_numpy(arr["z"]).to(DEV)
    video = m.decode(z, c["temporal_length"])
    assert list(video.shape) == list(arr["video"].shape)
    rel, psnr = _errors(video, torch.from_numpy(arr["video"]))
    print(f"fixture: rel {rel:.3e} psnr {psnr:.1f} dB")
    assert rel <= REL_TOL and psnr >= PSNR_MIN
    short = m.decode(z[:, :, :2].contiguous(), 5)
    rel, psnr = _errors(short, torch.from_numpy(arr["short"]))
    assert rel <= REL_TOL and psnr >= PSNR_MIN


@pytest.mark.parametrize("hidden,z_ch,mult,B,T,hw", [(64, 8, (1, 2, 4, 4), 1, 3, 8), (32, 16, (1, 2, 4, 4), 2, 2, 16),
                                                      (64, 4, (1, 2, 2, 4), 3, 1, 8)])
def test_decode_matches_oracle(hidden, z_ch, mult, B, T, hw):
    m, sd = _model(hidden, z_ch, z_ch, mult, seed=hidden + T)
    g = torch.Generator().manual_seed(5 * B + T)
    z = torch.randn((B, z_ch, T, hw, hw), generator=g)
    ref = VideoVAEDecoderOracle(sd, mult).decode(z)
    got = m.decode(z.to(DEV))
    assert got.shape == ref.shape == (B, 3, 1 + 4 * (T - 1), 8 * hw, 8 * hw)
    rel, psnr = _errors(got, ref)
    print(f"oracle case: rel {rel:.3e} psnr {psnr:.1f} dB")
    assert rel <= REL_TOL and psnr >= PSNR_MIN


def test_k600_shape_full_size_and_batch_independence():
    """K600 latents [16, 16, 16] x 5 tokens -> 17 frames of 128 x 128 (hidden 128, mult (1, 2, 4, 4)), the full-size
    decode: against the oracle's torch fp32 ops executed on the GPU (TF32 off), and samples of a batch do not interact."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd = _model(128, 16, 16, (1, 2, 4, 4), seed=3)
    g = torch.Generator().manual_seed(9)
    z = torch.randn((2, 16, 5, 16, 16), generator=g).to(DEV)
    full = m.decode(z, 17)
    assert full.shape == (2, 3, 17, 128, 128) and torch.isfinite(full).all()
    ref = VideoVAEDecoderOracle({k: v.to(DEV) for k, v in sd.items()}, (1, 2, 4, 4)).decode(z, 17)
    rel, psnr = _errors(full, ref)
    print(f"K600 shape: rel {rel:.3e} psnr {psnr:.1f} dB")
    assert rel <= REL_TOL and psnr >= PSNR_MIN
    # GroupNorm partial sums are split by a batch-dependent slab count (fp32 rounding differs in the last bit), so a
    # sample decoded alone equals its in-batch result up to the bf16 noise floor, not bit for bit
    one = m.decode(z[1:].contiguous(), 17)
    rel, psnr = _errors(one, full[1:])
    print(f"alone vs in batch: rel {rel:.3e} psnr {psnr:.1f} dB")
    assert rel <= REL_TOL and psnr >= PSNR_MIN + 10      # measured 8e-3 / 61 dB; the atomics make it vary run to run


def test_encode_and_cpu_are_refused():
    m = VideoVAE(hidden_size=32, hidden_size_mult=(1, 2, 2, 2))
    with pytest.raises(NotImplementedError):
        m.encode(torch.zeros(1))
    with pytest.raises(RuntimeError, match="CUDA only"):
        m.decode(torch.zeros(1, 4, 1, 4, 4))


def test_sample_all_videos_decodes_latents():
    """The caller's view (dfot_video.py:82-112): a latent-video configuration (temporal downsampling 4, spatial 8) samples
    5 latent tokens from 2 context tokens and returns 17 decoded frames in [0, 1]-space; `gt` comes from the batch when
    the dataset supplies it.  The decoded prediction equals the oracle's decode of the product's own latents."""
    from dfot_b200.algorithms.dfot.dfot_video import DFoTVideo
    from oracle.cases import algorithm_cfg
    cfg = algorithm_cfg(**{"backbone.hidden_size": 128, "backbone.depth": 2, "backbone.num_heads": 2,
                           "backbone.spatial_mlp_ratio": 4.0, "backbone.patch_size": 1, "x_shape": [3, 32, 32],
                           "latent.enabled": True, "latent.downsampling_factor": [4, 8], "latent.num_channels": 4,
                           "max_frames": 17, "n_frames": 17, "context_frames": 5, "diffusion.sampling_timesteps": 4,
                           "data_mean": [[[0.1]]] * 4, "data_std": [[[1.5]]] * 4})
    torch.manual_seed(0)
    algo = DFoTVideo(cfg)
    assert algo.x_shape == [4, 4, 4] and algo.n_tokens == 5 and algo.n_context_tokens == 2
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for _, p in algo.named_parameters():
            if bool((p == 0).all()):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    algo = algo.to(DEV).eval()
    latents = torch.randn((3, 5, 4, 4, 4), generator=g).to(DEV)
    batch = algo.on_after_batch_transfer({"latents": latents})
    out = algo._sample_all_videos(batch)                                   # no decoder configured: latents come back
    assert out["prediction"].shape == (3, 5, 4, 4, 4)
    vae, sd = _model(32, 4, 4, (1, 2, 2, 2), seed=11)
    algo.vae = vae
    torch.manual_seed(5)
    out = algo._sample_all_videos(batch)
    assert out["prediction"].shape == out["gt"].shape == (3, 17, 3, 32, 32)
    torch.manual_seed(5)
    algo.vae = None
    lat = algo._sample_all_videos(batch)["prediction"]                     # same seed -> same latents
    ref = VideoVAEDecoderOracle(sd, (1, 2, 2, 2)).decode(lat.cpu().permute(0, 2, 1, 3, 4), 17) * 0.5 + 0.5
    rel, psnr = _errors(out["prediction"].permute(0, 2, 1, 3, 4), ref)
    assert rel <= REL_TOL and psnr >= PSNR_MIN
    gt = torch.rand((3, 17, 3, 32, 32), device=DEV)
    algo.vae = vae
    batch = algo.on_after_batch_transfer({"latents": latents, "videos": gt})
    assert torch.equal(algo._sample_all_videos(batch)["gt"], gt)


# ---------------------------------------------------------------------------------------------- ImageVAE (2-D decoder)
def _image_model(dd, embed, seed):
    from dfot_b200.algorithms.vae import ImageVAE
    from oracle.image_vae import image_decoder_param_shapes, seeded_image_weights
    shapes = image_decoder_param_shapes(dd["ch"], dd["z_channels"], embed, tuple(dd["ch_mult"]), dd["num_res_blocks"])
    sd = seeded_image_weights(shapes, seed)
    m = ImageVAE(dict(ddconfig=dd, embed_dim=embed))
    m.load_state_dict(sd)
    return m.to(DEV), sd


def test_image_decode_matches_reference_fixture():
    with open(os.path.join(GOLDEN, "vae_image_decode.json")) as f:
        c = json.load(f)["case"]
    arr = dict(np.load(os.path.join(GOLDEN, "vae_image_decode.npz")))
    m, _ = _image_model(c["ddconfig"], c["embed_dim"], c["weight_seed"])
    images = m.decode(torch.from_numpy(arr["z"]).to(DEV))
    assert list(images.shape) == list(arr["images"].shape)
    rel, psnr = _errors(images, torch.from_numpy(arr["images"]))
    print(f"image fixture: rel {rel:.3e} psnr {psnr:.1f} dB")
    assert rel <= REL_TOL and psnr >= PSNR_MIN


@pytest.mark.parametrize("ch,mult,n,hw", [(128, (1, 2, 4, 4), 5, 8), (64, (1, 2), 3, 16), (32, (1, 2, 4, 4), 20, 4)])
def test_image_decode_matches_oracle(ch, mult, n, hw):
    """configurations/algorithm/image_vae.yaml's ddconfig (ch 128, mult 1-2-4-4, z 4) and two other topologies, with the
    oracle's torch fp32 ops executed on the GPU (TF32 off) for the wide one."""
    from oracle.image_vae import ImageVAEDecoderOracle
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dd = dict(double_z=True, z_channels=4, resolution=hw * 2 ** (len(mult) - 1), in_channels=3, out_ch=3, ch=ch,
              ch_mult=list(mult), num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    m, sd = _image_model(dd, 4, seed=ch + n)
    z = torch.randn((n, 4, hw, hw), generator=torch.Generator().manual_seed(n)).to(DEV)
    ref = ImageVAEDecoderOracle({k: v.to(DEV) for k, v in sd.items()}, mult).decode(z)
    got = m.decode(z)
    assert got.shape == ref.shape == (n, 3, hw * 2 ** (len(mult) - 1), hw * 2 ** (len(mult) - 1))
    rel, psnr = _errors(got, ref)
    print(f"image oracle case: rel {rel:.3e} psnr {psnr:.1f} dB")
    assert rel <= REL_TOL and psnr >= PSNR_MIN


def test_decode_of_image_latents_through_the_algorithm():
    """`_decode` of a [b, t, c, h, w] latent video with an image VAE (temporal downsampling 1): frames go through the
    decoder as a batch of images in chunks of vae.batch_size clips (base_pytorch_video_algo.py:555-629)."""
    from dfot_b200.algorithms.dfot.dfot_video import DFoTVideo
    from oracle.cases import algorithm_cfg
    from oracle.image_vae import ImageVAEDecoderOracle
    cfg = algorithm_cfg(**{"backbone.hidden_size": 64, "backbone.depth": 1, "backbone.num_heads": 1, "x_shape": [3, 64, 64],
                           "latent.enabled": True, "latent.downsampling_factor": [1, 8], "latent.num_channels": 4,
                           "max_frames": 4, "n_frames": 4, "vae.batch_size": 2})
    algo = DFoTVideo(cfg).to(DEV).eval()
    assert algo.x_shape == [4, 8, 8] and not algo.is_latent_video_vae
    dd = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    algo.vae, sd = _image_model(dd, 4, seed=2)
    lat = torch.randn((3, 4, 4, 8, 8), generator=torch.Generator().manual_seed(4)).to(DEV)
    frames = algo._decode(lat)
    assert frames.shape == (3, 4, 3, 64, 64)
    ref = ImageVAEDecoderOracle(sd, (1, 2, 2, 4)).decode(lat.cpu().reshape(12, 4, 8, 8)).reshape(3, 4, 3, 64, 64) * 0.5 + 0.5
    rel, psnr = _errors(frames, ref)
    assert rel <= REL_TOL and psnr >= PSNR_MIN
