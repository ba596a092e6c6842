"""Host-side helpers for the multi-GPU path (one process per GPU, torch.distributed).

The reference is single-device; everything here is new.  Two layouts:
  * data parallel (`BaseModel.enable_data_parallel`): parameters replicated, batch sharded, the loss
    normaliser n and all gradients summed over ranks => identical to one process on the global batch;
  * row-sharded item table (`shard_rows`): rank r owns rows [lo, hi) of E, its Adam state and gradient.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_rows(num_rows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced row ranges; the first `num_rows % world` ranks get one extra row."""
    base, extra = divmod(num_rows, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def owner_of(ids: torch.Tensor, num_rows: int, world: int) -> torch.Tensor:
    """Rank that owns each row id under `shard_rows` (vectorised, any device)."""
    base, extra = divmod(num_rows, world)
    cut = extra * (base + 1)                       # ids below `cut` live on the ranks with base+1 rows
    big = torch.div(ids, base + 1, rounding_mode='floor')
    small = extra + torch.div(ids - cut, max(base, 1), rounding_mode='floor')
    return torch.where(ids < cut, big, small)


def sum_over_ranks(tensors: Sequence[torch.Tensor], group=None) -> None:
    """In-place SUM all-reduce of each tensor (gradients, the valid-target count, the loss)."""
    if group is None and not dist.is_initialized():
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def split_batch(batch: dict, rank: int, world: int) -> dict:
    """Rows [rank::world] of every tensor of a global batch (what each data-parallel rank trains on)."""
    return {k: v[rank::world].contiguous() for k, v in batch.items()}
