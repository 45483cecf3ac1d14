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
