"""Data-parallel host logic: contiguous batch shards, one all-reduce(SUM) of the flat gradient.

Utterances are independent through the frontend, the forward and the per-sample loss; the only exchange step of the
path is the gradient mean.  Each rank scales its loss by 1 / global_batch (`loss_scale_batch` of the C ABI) so the
all-reduce is a plain SUM; BatchNorm statistics stay per rank (DDP semantics).
"""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global batch owned by `rank`; remainders go to the lowest ranks."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_flat_grads(grads: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM over ranks of the flat fp32 gradient buffer (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=group)
    return grads
