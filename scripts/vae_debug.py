"""Debug aid (GPU): stage-by-stage comparison of dfot_b200's VideoVAE.decode with the CPU oracle on a small case."""
import sys

import torch

sys.path.insert(0, ".")
from dfot_b200.algorithms.vae import VideoVAE  # noqa: E402
from dfot_b200.algorithms.vae.video_vae import PAD  # noqa: E402
from oracle.video_vae import VideoVAEDecoderOracle, decoder_param_shapes, seeded_weights  # noqa: E402

hidden, zc, mult, B, T, hw = (int(sys.argv[1]) if len(sys.argv) > 1 else 32), 4, (1, 2, 2, 2), 2, 3, 4
sd = seeded_weights(decoder_param_shapes(hidden, zc, zc, mult), 11)
m = VideoVAE(hidden_size=hidden, z_channels=zc, embed_dim=zc, hidden_size_mult=mult)
m.load_state_dict(sd)
m = m.cuda()
orc = VideoVAEDecoderOracle(sd, mult)
z = torch.randn((B, zc, T, hw, hw), generator=torch.Generator().manual_seed(1))
rec_o, rec_p = [], []


def wrap_o(name):
    f = getattr(orc, name)

    def g(*a):
        y = f(*a)
        if name != "conv" or a[0] in ("decoder.conv_in", "post_quant_conv", "decoder.conv_out"):
            rec_o.append((f"{name}:{a[0]}", y.clone()))
        return y
    setattr(orc, name, g)


for n in ("conv", "resblock", "attn", "upsample"):
    wrap_o(n)


def to_bcthw(t):
    return t[:, PAD:].permute(0, 4, 1, 2, 3).float().cpu()


for n in ("_resblock", "_attn"):
    f = getattr(m, n)

    def g(P, name, x, *a, _f=f, _n=n):
        y = _f(P, name, x, *a)
        rec_p.append((f"{_n}:{name}", to_bcthw(y)))
        return y
    setattr(m, n, g)
fc = m._conv


def gc(P, name, a16, out, *a, **k):
    fc(P, name, a16, out, *a, **k)
    if "upsample" in name or name in ("decoder.conv_in", "decoder.conv_out"):
        rec_p.append((f"_conv:{name}", to_bcthw(out)))


m._conv = gc
ref = orc.decode(z)
got = m.decode(z.cuda())
rec_o = [r for r in rec_o if not r[0].startswith("conv:post_quant")]
for (no, o), (np_, p) in zip(rec_o, rec_p):
    p = p[:, : o.shape[1]]
    print(f"{no:45s} {np_:45s} rel {((p - o).norm() / o.norm()).item():.3e}  max|ref| {o.abs().max().item():.3f}")
print("final rel", ((got.cpu() - ref).norm() / ref.norm()).item())
