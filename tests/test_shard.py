"""Host logic of the N > 1 path on CPU: world_size-2 gloo processes (127.0.0.1) exercising the sharding helpers bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from urban_road_filter_b200.shard import allreduce_max, allreduce_sum, seeds_for_rank, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 128, 257):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in shard_range(total, r, world)]
            assert got == list(range(total))
            sizes = [len(shard_range(total, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert not set(seeds_for_rank(128, 0)) & set(seeds_for_rank(128, 1))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_range(11, rank, world)
    mx = allreduce_max([10.0 + rank, 5.0 - rank])
    sm = allreduce_sum([len(mine), sum(mine)])
    dist.barrier()
    q.put((rank, list(mine), mx, sm))
    dist.destroy_process_group()


def test_two_rank_gloo_reductions():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] + res[1][1] == list(range(11))
    for _, _, mx, sm in res:
        assert mx == [11.0, 5.0]          # max over ranks of (10 + rank, 5 - rank)
        assert sm == [11, sum(range(11))]
