"""Sharding env instances over the GPUs of one box + the only cross-GPU exchange of the path.

Env instances are independent: rank ``r`` of ``G`` owns the contiguous
range of global env ids ``[offset_r, offset_r + count_r)``, network and attractor tables are
replicated, and the Philox counters use *global* env ids, so a sharded batch draws exactly what
the single-GPU batch would.  Nothing crosses GPUs on the step path; the episode statistics
(``missed`` / ``rew_recap`` / ``len_recap`` of bdq_model/__init__.py:169,179-231) are summed with
one small all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import _cabi

__all__ = ["TILE", "shard_range", "make_sharded_env", "allreduce_stats"]

TILE = 1024  # env_offset granularity required by the C-ABI (one kernel tile)


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """``(offset, count)`` of rank's shard: whole 1024-env tiles, as even as possible, the last
    (possibly partial) tile going to the last rank that owns any envs."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    tiles = (total_envs + TILE - 1) // TILE
    base, extra = divmod(tiles, world)
    first = rank * base + min(rank, extra)
    mine = base + (1 if rank < extra else 0)
    offset = first * TILE
    count = max(0, min(total_envs, (first + mine) * TILE) - offset)
    return offset, count


def make_sharded_env(network, total_envs: int, attractors=None, rank: Optional[int] = None,
                     world: Optional[int] = None, device=None, **kwargs):
    """This rank's :class:`VecPBNEnv` shard of a logical batch of ``total_envs`` instances."""
    from .vec_env import VecPBNEnv

    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    offset, count = shard_range(total_envs, rank, world)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return VecPBNEnv(network, count, attractors, device=device, env_offset=offset, **kwargs)


def allreduce_stats(stats: torch.Tensor, group=None, async_op: bool = False):
    """Sum the ``[PBN_N_STATS]`` int64 statistics vector over all ranks (in place).  Returns the
    work handle when ``async_op`` -- the caller can overlap it with the next steps."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def stats_dict(stats: torch.Tensor) -> Dict[str, int]:
    vals = stats.detach().cpu().tolist()
    return dict(zip(_cabi.STAT_NAMES[:7], vals[:7]))
