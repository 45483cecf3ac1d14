"""torchrun --nproc-per-node 2 scripts/check_decode_sharded.py — NCCL sanity of distributed.decode_sharded: 5 clips dealt
over the ranks (uneven split) decode to the same frames as one rank decoding all of them."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from dfot_b200 import distributed as D  # noqa: E402
from dfot_b200.algorithms.vae import VideoVAE  # noqa: E402

rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
vae = VideoVAE(hidden_size=64, z_channels=16, embed_dim=16, hidden_size_mult=(1, 2, 4, 4)).to(dev)
z = torch.randn((5, 16, 5, 16, 16), generator=torch.Generator().manual_seed(1)).to(dev)
out = D.decode_sharded(lambda lat: vae.decode(lat, 17), z)
ref = vae.decode(z, 17)
rel = ((out - ref).norm() / ref.norm()).item()
print(f"rank {dist.get_rank()}/{dist.get_world_size()}: decoded {tuple(out.shape)}, rel vs single-rank decode {rel:.3e}")
assert out.shape == ref.shape == (5, 3, 17, 128, 128) and rel <= 2e-2   # chunked vs whole-batch decode differ by the bf16 noise floor (GroupNorm partial-sum order)
dist.destroy_process_group()
