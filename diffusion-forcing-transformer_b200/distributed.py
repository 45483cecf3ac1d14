"""Multi-GPU sharding of the sampling path: one process per GPU, torch.distributed (NCCL over
NVLink on the B200 box, gloo in CPU tests).  The path shards along two independent axes and never along the
sequence:

  * samples  — every rank samples its own slice of the batch with no communication inside the loop; one
               all_gather of the finished samples at the end (`gather_samples`);
  * interpolation chunk batches — when there are fewer samples than dp shards, the chunk batches of an interpolation
               round (independent `_sample_sequence` calls) are dealt round-robin over the dp axis; every other shard
               replays the batch's noise draws only, and the owner broadcasts the finished batch
               (`broadcast_from_shard`), so the result equals the single-GPU rollout;
  * VAE decode — after the final gather every rank holds every sampled latent; the decode (per-sample independent) is
               dealt over ALL ranks and gathered once (`decode_sharded`);
  * history-guidance branches — within a branch group of `br` ranks (br divides nfe) each rank runs the backbone
               on its share of the branch rows of every sample; one all_gather of the backbone output per step,
               after which every member runs the identical fused K4 step (same noise seed), so x_t stays replicated.

Mesh: world = dp x br, rank = dp_index * br + br_index.
"""
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist


@dataclass
class BranchGroup:
    group: Optional[object]   # torch.distributed process group (None = default group)
    size: int
    rank: int                 # index of this process inside the group


@dataclass
class Mesh:
    world: int
    rank: int
    dp: int
    br: int
    dp_index: int
    br_index: int
    branch_group: Optional[BranchGroup]
    dp_group: Optional[object]


def build_mesh(br: int = 1) -> Mesh:
    """Create the dp x br mesh over the default process group (call after init_process_group)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world % br:
        raise ValueError(f"world size {world} is not divisible by the branch-group size {br}")
    dp = world // br
    dp_index, br_index = divmod(rank, br)
    branch_group, dp_group = None, None
    # new_group must be called by all ranks for every group, in the same order
    for d in range(dp):
        ranks = [d * br + j for j in range(br)]
        g = dist.new_group(ranks) if br > 1 else None
        if d == dp_index and br > 1:
            branch_group = BranchGroup(g, br, br_index)
    for j in range(br):
        ranks = [d * br + j for d in range(dp)]
        g = dist.new_group(ranks) if dp > 1 and br > 1 else None
        if j == br_index:
            dp_group = g
    return Mesh(world, rank, dp, br, dp_index, br_index, branch_group, dp_group)


def shard_batch(n: int, parts: int, index: int) -> slice:
    """Contiguous, balanced slice of a batch of n samples for shard `index` of `parts`."""
    base, extra = divmod(n, parts)
    start = index * base + min(index, extra)
    return slice(start, start + base + (1 if index < extra else 0))


def branch_rows(batch: int, nfe: int, bg: BranchGroup) -> List[int]:
    """Rows (b, j) of the (b h g)-ordered branch batch owned by this member: j ≡ rank (mod size)."""
    if nfe % bg.size:
        raise ValueError(f"nfe={nfe} is not divisible by the branch-group size {bg.size}")
    return [b * nfe + j for b in range(batch) for j in range(nfe) if j % bg.size == bg.rank]


def gather_branch_outputs(local: torch.Tensor, batch: int, nfe: int, bg: BranchGroup) -> torch.Tensor:
    """all_gather of the per-step backbone outputs inside the branch group, re-ordered to (b, j)."""
    per = nfe // bg.size
    parts = [torch.empty_like(local) for _ in range(bg.size)]
    dist.all_gather(parts, local.contiguous(), group=bg.group)
    # member m holds rows (b, j = m + size*k), k < per, ordered (b, k)
    stacked = torch.stack(parts, 0).reshape(bg.size, batch, per, *local.shape[1:])   # [m, b, k, ...]
    return stacked.permute(1, 2, 0, *range(3, stacked.ndim)).reshape(batch * nfe, *local.shape[1:]).contiguous()


def gather_samples(local: torch.Tensor, mesh: Mesh, counts: List[int]) -> torch.Tensor:
    """Final all_gather of the samples of all dp shards (ragged shards are padded to the largest)."""
    if mesh.dp == 1:
        return local
    biggest = max(counts)
    pad = torch.zeros((biggest, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(mesh.dp)]
    dist.all_gather(parts, pad, group=mesh.dp_group if mesh.br > 1 else None)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)


def broadcast_from_shard(t: torch.Tensor, owner_dp_index: int, mesh: Mesh) -> torch.Tensor:
    """Broadcast a finished chunk batch from dp shard `owner_dp_index` to the same branch member of every other shard
    (in place; x_t is replicated inside a branch group, so each member serves its own column of the mesh)."""
    if mesh.dp == 1:
        return t
    dist.broadcast(t, src=owner_dp_index * mesh.br + mesh.br_index, group=mesh.dp_group if mesh.br > 1 else None)
    return t


def decode_sharded(decode_fn, latents: torch.Tensor) -> torch.Tensor:
    """VAE decode of a batch every rank holds (the state after `gather_samples`): samples are independent, so rank r
    decodes samples [r * per, (r + 1) * per) (indices past the batch repeat the last sample so shapes agree) and one
    all_gather over the whole world returns the decoded batch to every rank.  No process group: plain `decode_fn`."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return decode_fn(latents)
    world, rank, n = dist.get_world_size(), dist.get_rank(), latents.shape[0]
    per = (n + world - 1) // world
    idx = torch.arange(rank * per, (rank + 1) * per, device=latents.device).clamp(max=n - 1)
    mine = decode_fn(latents.index_select(0, idx)).contiguous()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return torch.cat(parts, 0)[:n]
