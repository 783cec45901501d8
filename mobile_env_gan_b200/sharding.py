"""Multi-GPU: environments are independent (each reference MComCore owns all of its state,
base.py:69-79), so rank r simply owns a contiguous slice of the global env range and the step
path has NO collective.  Philox counters use the global env id (``env_offset``), so results do
not depend on the number of ranks.  The only exchange is the optional all-gather of per-env
episode statistics."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_envs(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(env_offset, num_envs) of ``rank``: contiguous, sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(total_envs, world_size)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def sharded_config(total_envs: int, rank: int = None, world_size: int = None) -> dict:
    """Config keys for this rank (reads torch.distributed when rank/world_size are omitted)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    offset, count = shard_envs(total_envs, rank, world_size)
    return {"env_offset": offset, "num_envs": count}


def gather_episode_stats(stats: torch.Tensor, total_envs: int, group=None) -> torch.Tensor:
    """All-gathers per-env statistics [E_local, K] into [total_envs, K] on every rank
    (NCCL over NVLink for CUDA tensors, gloo on the CPU).  Off the step path."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return stats
    world = dist.get_world_size(group)
    counts = [shard_envs(total_envs, r, world)[1] for r in range(world)]
    pad = max(counts)
    buf = stats.new_zeros((pad,) + tuple(stats.shape[1:]))
    buf[: stats.shape[0]] = stats
    out = stats.new_empty((world * pad,) + tuple(stats.shape[1:]))
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = [out[r * pad: r * pad + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0)
