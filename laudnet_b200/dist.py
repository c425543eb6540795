"""Batch-sharded inference across the GPUs of one box.

Images are independent in eval mode (BatchNorm uses running statistics, every
gate is per sample), so the path shards with NO data-path collective: rank r
runs the hot path on a contiguous slice of the batch with replicated weights.
The one exchange step is the gather of the per-rank logits
`[B/N, n_cls] -> [B, n_cls]` (plus, optionally, the per-block active counts so
the reference's batch-mean densities can be reproduced globally) - a single
`all_gather_into_tensor` over NCCL/NVLink on the GPU box, `gloo` in CPU tests.

The reference has no equivalent (its only collectives are DDP gradient
all-reduces and scalar metric all-reduces, train/main.py:326,665-697).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of a batch for `rank`; the first `batch % world`
    ranks get one extra sample (ragged batches are allowed)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_shard(batch: int, world: int) -> int:
    return (batch + world - 1) // world


def allgather_logits(local: torch.Tensor, batch: int, group: Optional[dist.ProcessGroup] = None,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Gather row-sharded logits.  `local` is this rank's [b_r, n_cls] slice
    (b_r from `shard_range`); returns [batch, n_cls] on every rank.

    One collective: ranks pad their slice to the common `max_shard` rows so a
    single `all_gather_into_tensor` moves everything; the pad rows are dropped
    when the result is compacted (only when the batch is ragged).
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local if out is None else out.copy_(local)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(batch, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: expected {hi - lo} local rows, got {local.shape[0]}")
    rows = max_shard(batch, world)
    n = local.shape[1]
    if local.shape[0] == rows:
        send = local.contiguous()
    else:
        send = local.new_zeros((rows, n))
        send[: local.shape[0]] = local
    gathered = local.new_empty((world * rows, n))
    dist.all_gather_into_tensor(gathered, send, group=group)
    if batch == world * rows:
        return gathered if out is None else out.copy_(gathered)
    res = out if out is not None else local.new_empty((batch, n))
    for r in range(world):
        a, b = shard_range(batch, r, world)
        res[a:b] = gathered[r * rows: r * rows + (b - a)]
    return res


def allreduce_counts(counts: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Sum the per-block active counts [n_blocks, 4] (int32) over ranks, so the
    densities / flops statistics equal the unsharded batch's (they are batch
    means in the reference, laud_resnet.py:121-147).  Not on the logits path."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


class ShardedClassifier:
    """model(x_shard) on every rank + one logits all-gather.

    `model` is a `laudnet_b200.ResNet` (or anything with `forward_logits`);
    `images_fn(lo, hi)` is not needed - callers pass their own shard.
    """

    def __init__(self, model, group: Optional[dist.ProcessGroup] = None):
        self.model = model
        self.group = group

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if dist.is_initialized() else 0

    def my_range(self, batch: int) -> Tuple[int, int]:
        return shard_range(batch, self.rank, self.world)

    def __call__(self, x_shard: torch.Tensor, batch: int) -> torch.Tensor:
        local = self.model.forward_logits(x_shard)
        return allgather_logits(local, batch, self.group)
