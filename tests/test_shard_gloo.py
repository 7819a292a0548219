"""N>1 host logic on CPU: world_size-2 gloo run of the pair sharding and the
throughput aggregation used by bench.py (no GPU, no kernels)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crossscalepatchmatch_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.pairs_for_rank(5, world, rank)
        units = 100.0 * len(mine)            # pretend every pair is 100 units
        secs = [1.0 + rank, 2.0 - rank]       # rank-dependent timings
        dist.barrier()
        u, t = shard.aggregate(units, secs)
        q.put((rank, mine, u, t))
    finally:
        dist.destroy_process_group()


def test_pairs_for_rank():
    assert shard.pairs_for_rank(8, 8, 3) == [3]
    assert shard.pairs_for_rank(5, 2, 0) == [0, 2, 4] and shard.pairs_for_rank(5, 2, 1) == [1, 3]
    got = sorted(sum((shard.pairs_for_rank(11, 4, r) for r in range(4)), []))
    assert got == list(range(11))
    with pytest.raises(ValueError):
        shard.pairs_for_rank(4, 2, 2)


def test_aggregate_single_process():
    assert shard.aggregate(7.0, [0.5, 0.25]) == (7.0, [0.5, 0.25])


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    for _, _, u, t in res:
        assert u == 500.0            # sum over ranks
        assert t == [2.0, 2.0]       # max over ranks, per entry
