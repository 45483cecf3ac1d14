"""Multi-GPU sharding of the sampling path: one process per GPU, torch.distributed (NCCL over
NVLink on the B200 box, gloo in CPU tests).  The path shards along two independent axes and never along the
sequence:

  * samples  — every rank samples its own slice of the batch with no communication inside the loop; one
               all_gather of the finished samples at the end (`gather_samples`);
  * forward rows — when there are fewer samples than dp shards (the single-sample 200-frame rollout), the sampler
               state (x_t, plans, noise) is REPLICATED on every rank and only the backbone forward is sharded: the branch
               rows of a keyframe window, or of ALL chunk batches of an interpolation round advancing in lockstep, are dealt
               in equal contiguous blocks over the whole world (`RowShard`), one `all_gather_into_tensor` of the backbone
               outputs per step, then every rank runs the identical fused K4 steps (same noise seed);
  * VAE decode — after the final gather every rank holds every sampled latent; the decode (per-sample independent) is
               dealt over ALL ranks and gathered once (`decode_sharded`);
  * history-guidance branches — within a branch group of `br` ranks each rank runs the backbone on its contiguous
               block of the group's (sample, branch) rows (`RowShard` over the group: with one sample per group that is one
               branch per member); one all_gather_into_tensor of the backbone output per step, after which every member
               runs the identical fused K4 step (same noise seed), so x_t stays replicated inside the group.

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


class RowShard:
    """Deal n forward-rows over the whole world in equal contiguous blocks of `per = ceil(n / world)` rows (the last
    blocks may be short or empty) and bring the per-row outputs back to every rank with ONE all_gather_into_tensor into a
    [world * per, ...] buffer whose first n rows are the outputs in row order — no list gather, no re-ordering copies."""

    def __init__(self, world: Optional[int] = None, rank: Optional[int] = None, group=None):
        self.world = dist.get_world_size(group) if world is None else world
        self.rank = dist.get_rank(group) if rank is None else rank
        self.group = group
        self._bufs = {}

    def block(self, n: int):
        """(per, start, stop): this rank forwards rows [start, stop) of the n."""
        per = -(-n // self.world)
        start = min(self.rank * per, n)
        return per, start, min(start + per, n)

    def gather(self, local: Optional[torch.Tensor], n: int, row_shape, dtype, device) -> torch.Tensor:
        """local: [stop - start, *row_shape] outputs of this rank's rows (None when it has none) -> [n, *row_shape]."""
        if self.world == 1:
            return local
        per, start, stop = self.block(n)
        key = (n, tuple(row_shape), dtype, str(device))
        buf = self._bufs.get(key)
        if buf is None:
            buf = self._bufs[key] = (torch.empty((self.world * per, *row_shape), dtype=dtype, device=device),
                                     torch.zeros((per, *row_shape), dtype=dtype, device=device))
        full, mine = buf
        if stop > start:
            mine[: stop - start].copy_(local)
        dist.all_gather_into_tensor(full, mine, group=self.group)
        return full[:n]


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
