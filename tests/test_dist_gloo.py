"""Multi-rank host logic on CPU: shard ranges and the statistics all-reduce over gloo (world_size 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pbn_rl_b200.dist import TILE, allreduce_stats, shard_range


@pytest.mark.parametrize("total", [0, 1, 1023, 1024, 1025, 4096, 1 << 20, (1 << 20) + 5, 70000])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_ranges_tile_the_batch(total, world):
    covered = 0
    for r in range(world):
        off, cnt = shard_range(total, r, world)
        assert off % TILE == 0
        assert off == covered or cnt == 0
        covered += cnt
    assert covered == total
    sizes = [shard_range(total, r, world)[1] for r in range(world)]
    assert max(sizes) - min(sizes) < 2 * TILE or total < world * TILE


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, cnt = shard_range(total, rank, world)
    # per-rank "episode statistics": steps = shard size, episodes = rank-dependent
    stats = torch.tensor([cnt, 10 * (rank + 1), 3, 7, 40 + rank, 5, 1, 0], dtype=torch.int64)
    allreduce_stats(stats)
    ranges = [None] * world
    dist.all_gather_object(ranges, (off, cnt))
    if rank == 0:
        out.put((stats.tolist(), ranges))
    dist.destroy_process_group()


def test_stats_allreduce_world2():
    total = 5 * TILE + 17
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats, ranges = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert stats[0] == total                      # steps add up to the logical batch
    assert stats[1] == 10 + 20 and stats[4] == 40 + 41
    assert ranges[0][0] == 0 and ranges[1][0] == ranges[0][1]
    assert sum(c for _, c in ranges) == total


def test_parse_cpulist_and_numa_binding_is_harmless(tmp_path):
    from pbn_rl_b200.dist import bind_to_gpu_numa, parse_cpulist
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("") == []
    # a fake two-node sysfs tree: rank 1 of 2 lands on node 1 when the GPU's node is unknown (no CUDA device here)
    for node, cpus in ((0, "0"), (1, "0")):
        d = tmp_path / "devices" / "system" / "node" / ("node%d" % node)
        d.mkdir(parents=True)
        (d / "cpulist").write_text(cpus + "\n")
    before = os.sched_getaffinity(0)
    info = bind_to_gpu_numa(0, rank=1, world=2, sysfs=str(tmp_path))
    assert info["node"] in (None, 1)
    os.sched_setaffinity(0, before)
