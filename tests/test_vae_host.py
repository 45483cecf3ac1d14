"""No-GPU: host logic of the VAE-decode row — the product's VideoVAE has the reference's decode-side state-dict keys,
loads reference-style checkpoints (raw and EMA), refuses what is out of scope, and BaseVideoAlgo._run_vae follows the
reference's layout / chunking rules (base_pytorch_video_algo.py:555-585)."""
import json
import os

import pytest
import torch

from dfot_b200.algorithms.vae import VideoVAE
from helpers import GOLDEN
from oracle.video_vae import decoder_param_shapes, seeded_weights


def _meta():
    with open(os.path.join(GOLDEN, "vae_video_decode.json")) as f:
        return json.load(f)


def test_state_dict_keys_and_shapes_match_reference():
    meta = _meta()
    c = meta["case"]
    m = VideoVAE(hidden_size=c["hidden_size"], z_channels=c["z_channels"], embed_dim=c["embed_dim"],
                 hidden_size_mult=tuple(c["hidden_size_mult"]), resolution=c["resolution"],
                 temporal_length=c["temporal_length"])
    assert list(m.state_dict().keys()) == meta["keys"]
    shapes = dict(decoder_param_shapes(c["hidden_size"], c["z_channels"], c["embed_dim"], tuple(c["hidden_size_mult"])))
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == shapes


@pytest.mark.parametrize("ema", [False, True])
def test_from_pretrained_reads_reference_checkpoints(tmp_path, ema):
    cfg = dict(hidden_size=32, z_channels=4, embed_dim=4, hidden_size_mult=[1, 2, 2, 2], resolution=32, temporal_length=9)
    sd = seeded_weights(decoder_param_shapes(32, 4, 4, (1, 2, 2, 2)), 5)
    # a full-model checkpoint: encoder tensors first, Lightning prefix `vae.`, a loss module beside it
    full = {"vae.encoder.conv_in.weight": torch.zeros(32, 3, 3, 3), "vae.encoder.conv_in.bias": torch.zeros(32)}
    full.update({f"vae.{k}": v for k, v in sd.items()})
    full["loss.discriminator.w"] = torch.zeros(3)
    ckpt = {"model_cfg": cfg, "optimizer_states": [], "state_dict": full}
    if ema:
        ckpt["state_dict"] = {k: torch.full_like(v, 7.0) for k, v in full.items()}      # raw weights must be ignored
        ckpt["optimizer_states"] = [{"ema": [v for k, v in full.items() if k.startswith("vae.")]}]
    path = str(tmp_path / "vae.ckpt")
    torch.save(ckpt, path)
    m = VideoVAE.from_pretrained(path)
    got = m.state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd) and len(got) == len(sd)


def test_out_of_scope_is_refused():
    with pytest.raises(NotImplementedError):
        VideoVAE(is_causal=False)
    with pytest.raises(NotImplementedError):
        VideoVAE(decoder_attention="AttnBlock3DFix")
    m = VideoVAE(hidden_size=32, hidden_size_mult=(1, 2, 2, 2), encoder_conv_in="Conv2d")   # encoder knobs are ignored
    with pytest.raises(NotImplementedError):
        m.encode(torch.zeros(1))
    with pytest.raises(RuntimeError, match="CUDA only"):
        m.decode(torch.zeros(1, 4, 1, 4, 4))
    with pytest.raises(RuntimeError):
        m.load_state_dict({"decoder.conv_in.conv.weight": torch.zeros(1)})


def test_run_vae_layout_and_chunking():
    from dfot_b200.algorithms.common.base_pytorch_video_algo import BaseVideoAlgo
    from dfot_b200.config import to_config

    class Stub(BaseVideoAlgo):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.cfg = to_config({"vae": {"batch_size": 2}})
            self.temporal_downsampling_factor = 4
            self.is_latent_video_vae = True
            self.vae = None
            self.calls = []

    algo = Stub()

    def fake_decode(y):                     # b c t h w -> b 3 (4(t-1)+1) 2h 2w
        algo.calls.append(tuple(y.shape))
        b, c, t, h, w = y.shape
        return y[:, :3].repeat_interleave(4, 2)[:, :, : 4 * (t - 1) + 1].repeat_interleave(2, 3).repeat_interleave(2, 4)

    x = torch.arange(5 * 3 * 4 * 2 * 2, dtype=torch.float32).reshape(5, 3, 4, 2, 2)         # b t c h w
    out = algo._run_vae(x, "b t c h w", fake_decode)
    assert out.shape == (5, 9, 3, 4, 4)
    assert algo.calls == [(2, 4, 3, 2, 2), (2, 4, 3, 2, 2), (1, 4, 3, 2, 2)]                 # chunks of vae.batch_size
    assert torch.equal(out[:, 0, :, ::2, ::2], x[:, 0, :3])
    with pytest.raises(NotImplementedError):
        algo._encode(x)


def test_run_vae_feeds_image_vaes_frame_by_frame():
    from dfot_b200.algorithms.common.base_pytorch_video_algo import BaseVideoAlgo
    from dfot_b200.config import to_config

    class Stub(BaseVideoAlgo):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.cfg = to_config({"vae": {"batch_size": 2}})
            self.temporal_downsampling_factor = 1
            self.is_latent_video_vae = False
            self.vae = None
            self.calls = []

    algo = Stub()

    def fake_decode(y):                     # (b t) c h w -> (b t) 3 2h 2w
        algo.calls.append(tuple(y.shape))
        return y[:, :3].repeat_interleave(2, 2).repeat_interleave(2, 3)

    x = torch.arange(3 * 4 * 5 * 2 * 2, dtype=torch.float32).reshape(3, 4, 5, 2, 2)         # b t c h w
    out = algo._run_vae(x, "b t c h w", fake_decode)
    assert out.shape == (3, 4, 3, 4, 4)
    assert algo.calls == [(8, 5, 2, 2), (4, 5, 2, 2)]
    assert torch.equal(out[:, :, :, ::2, ::2], x[:, :, :3])


def test_image_vae_keys_checkpoint_and_refusals(tmp_path):
    from dfot_b200.algorithms.vae import ImageVAE
    from oracle.image_vae import image_decoder_param_shapes, seeded_image_weights
    with open(os.path.join(GOLDEN, "vae_image_decode.json")) as f:
        meta = json.load(f)
    c = meta["case"]
    dd = c["ddconfig"]
    m = ImageVAE(dict(ddconfig=dd, embed_dim=c["embed_dim"]))
    assert list(m.state_dict().keys()) == meta["keys"]
    shapes = image_decoder_param_shapes(dd["ch"], dd["z_channels"], c["embed_dim"], tuple(dd["ch_mult"]), dd["num_res_blocks"])
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(shapes)
    sd = seeded_image_weights(shapes, 3)
    full = {"encoder.conv_in.weight": torch.zeros(32, 3, 3, 3), "loss.logvar": torch.zeros(())}
    full.update(sd)
    path = str(tmp_path / "image_vae.ckpt")
    torch.save({"cfg": dict(ddconfig=dd, embed_dim=c["embed_dim"]), "state_dict": full}, path)
    got = ImageVAE.from_pretrained(path).state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd) and len(got) == len(sd)
    with pytest.raises(NotImplementedError):
        ImageVAE.from_pretrained("diffuser:madebyollin/sdxl-vae-fp16-fix")
    with pytest.raises(NotImplementedError):
        ImageVAE(dict(ddconfig={**dd, "attn_resolutions": [16]}, embed_dim=4))
    with pytest.raises(NotImplementedError):
        m.encode(torch.zeros(1))
    with pytest.raises(RuntimeError, match="CUDA only"):
        m.decode(torch.zeros(1, 4, 8, 8))


def test_weights_pack_into_kernel_layout():
    """[Cout, kt, 3, 3, Cin] bf16 conv weights (Cin of conv_in / post_quant padded to one K tile, Cout of conv_out to 8),
    [Cout, Cin] for 1x1 convolutions, fp32 biases and norm affine pairs — for both decoders."""
    from dfot_b200.algorithms.vae import ImageVAE
    v = VideoVAE(hidden_size=32, z_channels=4, embed_dim=4, hidden_size_mult=(1, 2, 2, 2))
    P = v._pack(torch.device("cpu"))
    assert tuple(P["decoder.conv_in"][0].shape) == (64, 3, 3, 3, 64) and P["decoder.conv_in"][0].dtype == torch.bfloat16
    assert tuple(P["post_quant_conv"][0].shape) == (64, 64) and tuple(P["decoder.conv_out"][0].shape) == (8, 3, 3, 3, 32)
    assert tuple(P["decoder.up.1.upsample.conv"][0].shape) == (64, 1, 3, 3, 64)
    assert tuple(P["decoder.up.2.upsample.conv"][0].shape) == (64, 3, 3, 3, 64)
    assert tuple(P["decoder.mid.attn_1.q"][0].shape) == (64, 64) and tuple(P["decoder.norm_out"][0].shape) == (32,)
    w = v.state_dict()["decoder.mid.block_1.conv1.conv.weight"]
    assert torch.equal(P["decoder.mid.block_1.conv1"][0], w.permute(0, 2, 3, 4, 1).to(torch.bfloat16))
    dd = dict(double_z=True, z_channels=4, resolution=32, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    m = ImageVAE(dict(ddconfig=dd, embed_dim=4))
    P = m._pack(torch.device("cpu"))
    assert tuple(P["decoder.conv_in"][0].shape) == (64, 1, 3, 3, 64) and tuple(P["decoder.conv_out"][0].shape) == (8, 1, 3, 3, 32)
    assert tuple(P["decoder.up.1.upsample.conv"][0].shape) == (64, 1, 3, 3, 64) and tuple(P["post_quant_conv"][0].shape) == (64, 64)
    assert tuple(P["decoder.up.0.block.0.nin_shortcut"][0].shape) == (32, 64) and "decoder.up.1.block.0.nin_shortcut" not in P
    w = m.state_dict()["decoder.mid.block_2.conv2.weight"]
    assert torch.equal(P["decoder.mid.block_2.conv2"][0][:, 0], w.permute(0, 2, 3, 1).to(torch.bfloat16))
    n_convs = sum(1 for k, t in m.state_dict().items() if t.ndim == 4)
    n_norms = sum(1 for k, t in m.state_dict().items() if t.ndim == 1 and k.endswith(".weight"))
    assert len(P) - 1 == n_convs + n_norms


def test_decoder_host_orchestration_with_emulated_kernels(monkeypatch):
    """Both decoders' host side (weight packing, padded-clip buffer flow, pointer offsets of the causal convolutions,
    strided GroupNorm, per-frame attention GEMMs, upsampling schedule) driven on the CPU by contract emulations of the
    kernels (tests/ops_emulation.py, bf16 operand rounding included) against the fixtures of the executed reference."""
    import numpy as np
    import ops_emulation
    from dfot_b200.algorithms.vae import ImageVAE
    from oracle.image_vae import image_decoder_param_shapes, seeded_image_weights
    ops_emulation.install(monkeypatch)

    def close(got, ref):
        got, ref = got.float(), torch.from_numpy(ref)
        assert list(got.shape) == list(ref.shape)
        assert ((got - ref).norm() / ref.norm()).item() <= 2e-2

    c = _meta()["case"]
    arr = dict(np.load(os.path.join(GOLDEN, "vae_video_decode.npz")))
    v = VideoVAE(hidden_size=c["hidden_size"], z_channels=c["z_channels"], embed_dim=c["embed_dim"],
                 hidden_size_mult=tuple(c["hidden_size_mult"]))
    v.load_state_dict(seeded_weights(decoder_param_shapes(c["hidden_size"], c["z_channels"], c["embed_dim"],
                                                          tuple(c["hidden_size_mult"])), c["weight_seed"]))
    z = torch.from_numpy(arr["z"])
    close(v.decode(z, c["temporal_length"]), arr["video"])
    close(v.decode(z[:, :, :2].contiguous(), 5), arr["short"])

    with open(os.path.join(GOLDEN, "vae_image_decode.json")) as f:
        ci = json.load(f)["case"]
    dd = ci["ddconfig"]
    arr = dict(np.load(os.path.join(GOLDEN, "vae_image_decode.npz")))
    m = ImageVAE(dict(ddconfig=dd, embed_dim=ci["embed_dim"]))
    m.load_state_dict(seeded_image_weights(image_decoder_param_shapes(dd["ch"], dd["z_channels"], ci["embed_dim"],
                                                                      tuple(dd["ch_mult"]), dd["num_res_blocks"]), ci["weight_seed"]))
    close(m.decode(torch.from_numpy(arr["z"])), arr["images"])
