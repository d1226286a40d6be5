"""Environment sharding across the GPUs of one box.

Environments are independent, so the step path needs no collective: rank g owns the contiguous
range [g*N/G, (g+1)*N/G) and keeps its state resident on its own device for the whole rollout.
The only exchange is an optional end-of-rollout reduction of a few diagnostic scalars
(sum KE, sum PE, sum spring energy, flagged environments) with torch.distributed
(NCCL over NVLink on the GPU box; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous environment range [lo, hi) of `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return n_total * rank // world, n_total * (rank + 1) // world


def shard_sizes(n_total: int, world: int):
    return [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]


def reduce_diagnostics(sums, group=None):
    """All-reduce (sum) the 4-vector written by MechanismState.energy_sums_device / a CPU tensor of
    the same layout. No-op without an initialised process group. Returns the tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums
