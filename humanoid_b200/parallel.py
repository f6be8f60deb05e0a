"""Multi-GPU host logic: env partitioning and the one collective on the path.

The step shards by env with no cross-GPU traffic; the only exchange is the
RunningNorm statistics all-reduce (PHC/policies/running_norm.py:23-34 computed over the
concatenated batch).  Pure torch.distributed plumbing: works with NCCL on GPUs and gloo on CPU.
"""

from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def env_partition(num_envs: int, world_size: int, rank: int, multiple: int = 8) -> Tuple[int, int]:
    """Contiguous env range [lo, hi) owned by ``rank``.  Boundaries fall on multiples of
    ``multiple`` (one kernel block) while they can, the last rank takes the remainder."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    blocks = -(-num_envs // multiple)
    per, extra = divmod(blocks, world_size)
    lo_b = rank * per + min(rank, extra)
    hi_b = lo_b + per + (1 if rank < extra else 0)
    return min(lo_b * multiple, num_envs), min(hi_b * multiple, num_envs)


def reduce_moments(sums: torch.Tensor, rows: int, group=None) -> torch.Tensor:
    """[sum(C) | sum of squares(C)] fp64 partials + local row count -> one all-reduce (SUM).
    Returns the payload ``[2C + 1]`` (last entry = total rows); identical on every rank."""
    if sums.dtype != torch.float64:
        raise TypeError("moment partials must be float64")
    payload = torch.cat([sums, torch.tensor([float(rows)], dtype=torch.float64, device=sums.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(payload, op=dist.ReduceOp.SUM, group=group)
    return payload
