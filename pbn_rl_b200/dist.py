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

__all__ = ["TILE", "shard_range", "make_sharded_env", "allreduce_stats", "parse_cpulist", "gpu_numa_node", "bind_to_gpu_numa"]

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


# ---- host placement for the host-buffer path (pbn_step_host): the rank's thread and its page-locked buffers
# ---- belong on the NUMA node the GPU hangs off -- every rank pinning on node 0 shares one memory controller ----

def parse_cpulist(text: str):
    """``"0-3,8,10-11"`` (sysfs cpulist syntax) -> sorted list of CPU ids."""
    cpus = []
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-", 1)
            cpus.extend(range(int(lo), int(hi) + 1))
        else:
            cpus.append(int(part))
    return sorted(set(cpus))


def gpu_numa_node(device_index: int, sysfs: str = "/sys") -> Optional[int]:
    """NUMA node of the GPU's PCIe function (sysfs ``numa_node``), or None if unknown / not a NUMA machine."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("%s/bus/pci/devices/%s/numa_node" % (sysfs, bus)) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa(device_index: int, rank: int = 0, world: int = 1, sysfs: str = "/sys") -> Dict[str, object]:
    """Pin the calling thread to the CPUs of the GPU's NUMA node (page-locked buffers allocated afterwards are then
    first-touched there).  When sysfs reports no node for the GPU but the machine has several nodes, the ranks are
    spread round-robin over the nodes instead, so that they do not all pin their buffers behind one memory
    controller.  Returns what was done (reported by bench.py); never raises."""
    import os
    info: Dict[str, object] = {"node": None, "cpus": None, "how": "unchanged"}
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("%s/devices/system/node" % sysfs) if d.startswith("node") and d[4:].isdigit())
        if len(nodes) < 2:
            info["how"] = "single NUMA node"
            return info
        node = gpu_numa_node(device_index, sysfs)
        how = "GPU's own node (sysfs numa_node)"
        if node is None or node not in nodes:
            node = nodes[rank % len(nodes)]
            how = "round-robin over %d nodes (GPU node unknown)" % len(nodes)
        with open("%s/devices/system/node/node%d/cpulist" % (sysfs, node)) as f:
            cpus = parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(node=node, cpus=len(cpus), how=how)
    except Exception as exc:  # placement is an optimisation, never a failure
        info["how"] = "unchanged (%s)" % type(exc).__name__
    return info
