"""DC-AE decoder (DMLab / Minecraft latents) without a GPU: the oracle against the fixture of the
executed reference, and the product's host side (state-dict layout, BatchNorm folding, [q|k|v] group layout, pixel-shuffle
shortcuts, buffer flow) on the CPU contract emulations of the kernels against the same fixture."""
import json
import os

import numpy as np
import pytest
import torch

import ops_emulation
from dfot_b200.algorithms.vae import MyAutoencoderDC
from helpers import GOLDEN
from oracle.dc_ae import DCAEDecoderOracle, decoder_param_shapes, seeded_weights


def _fixture():
    with open(os.path.join(GOLDEN, "dcae_decode.json")) as f:
        meta = json.load(f)
    arr = np.load(os.path.join(GOLDEN, "dcae_decode.npz"))
    sd = seeded_weights(decoder_param_shapes(meta["cfg"]), meta["weight_seed"])
    return meta, torch.from_numpy(arr["z"]), torch.from_numpy(arr["image"]), sd


def test_oracle_matches_the_executed_reference():
    meta, z, image, sd = _fixture()
    out = DCAEDecoderOracle(sd, meta["cfg"]).decode(z)
    assert out.shape == image.shape == (3, 3, 64, 64)
    assert (out - image).abs().max().item() <= 2e-5


def test_state_dict_layout_is_the_reference_s():
    meta, _, _, sd = _fixture()
    model = MyAutoencoderDC(meta["cfg"])
    assert list(model.state_dict().keys()) == meta["keys"]                       # names AND registration order
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    model.load_state_dict({**sd, "encoder.conv_in.conv.weight": torch.zeros(1)})  # encoder tensors are ignored
    with pytest.raises(RuntimeError):
        model.load_state_dict({k: v for k, v in sd.items() if "norm_out" not in k})
    with pytest.raises(NotImplementedError):
        model.encode(torch.zeros(1, 3, 64, 64))


def test_host_orchestration_on_emulated_kernels(monkeypatch):
    meta, z, image, sd = _fixture()
    ops_emulation.install(monkeypatch)
    model = MyAutoencoderDC(meta["cfg"])
    model.load_state_dict(sd)
    out = model.decode(z)
    assert out.shape == image.shape
    rel = ((out - image).norm() / image.norm()).item()
    assert rel <= 2e-2, rel                                                      # bf16 conv / GEMM operands (emulated)
    mse = ((out - image) ** 2).mean().item()
    rng = (image.max() - image.min()).item()
    assert 10 * np.log10(rng * rng / mse) >= 40.0


def test_load_vae_selects_dc_ae(tmp_path, monkeypatch):
    """`vae.name: dc_ae_preprocessor` (dmlab_video_generation.yaml) -> MyAutoencoderDC.from_pretrained(cfg=vae cfg) and
    `_decode` = vae.decode(y) * 0.5 + 0.5 over the frames of a clip (base_pytorch_video_algo.py:511-520, 599-605)."""
    from oracle.cases import algorithm_cfg
    meta, z, image, sd = _fixture()
    ckpt = str(tmp_path / "dcae.pth")
    torch.save(sd, ckpt)
    vae_cfg = dict(meta["cfg"], name="dc_ae_preprocessor", pretrained_path=ckpt, pretrained_kwargs={}, use_fp16=False,
                   batch_size=2)
    cfg = algorithm_cfg(**{"backbone.hidden_size": 64, "backbone.depth": 1, "backbone.num_heads": 1, "x_shape": [3, 64, 64],
                           "latent.enabled": True, "latent.downsampling_factor": [1, 8], "latent.num_channels": 8,
                           "max_frames": 3, "n_frames": 3, "context_frames": 1, "vae": vae_cfg,
                           "data_mean": [[[0.0]]] * 8, "data_std": [[[1.0]]] * 8})
    ops_emulation.install(monkeypatch)
    from dfot_b200.algorithms.dfot import DFoTVideo
    algo = DFoTVideo(cfg)
    frames = algo._decode(z.unsqueeze(0))                                        # one clip of 3 frames: [1, 3, 8, 8, 8]
    assert type(algo.vae).__name__ == "MyAutoencoderDC" and frames.shape == (1, 3, 3, 64, 64)
    ref = image.unsqueeze(0) * 0.5 + 0.5
    assert ((frames - ref).norm() / ref.norm()).item() <= 2e-2
