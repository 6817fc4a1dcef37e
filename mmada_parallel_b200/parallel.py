"""Multi-GPU layout of the denoising path: independent prompt replicas, one process per GPU, no data-path collective
(every generate_ti2ti call owns its ids and RNG). torch.distributed is used only for the barrier and
for reducing the timing/throughput scalars (max over ranks), never on the data path."""
from __future__ import annotations

from typing import List, Sequence, TypeVar

import torch
import torch.distributed as dist

T = TypeVar("T")


def shard_prompts(prompts: Sequence[T], rank: int, world_size: int) -> List[T]:
    """Round-robin assignment of independent prompts to ranks (rank r takes r, r+W, r+2W, ...)."""
    return [p for i, p in enumerate(prompts) if i % world_size == rank]


def _reduce(value: float, op, device) -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(value: float, device="cuda") -> float:
    return _reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value: float, device="cuda") -> float:
    return _reduce(value, dist.ReduceOp.SUM, device)
