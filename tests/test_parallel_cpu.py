"""The N>1 host logic (game sharding, max/sum reduction, variable-length sample gather) on CPU: world_size 2, gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alphagpu_b200 import parallel


def test_shard_games_partition():
    for total, world in [(32768, 8), (1000, 3), (7, 8), (65536, 2)]:
        seen = []
        for r in range(world):
            base, cnt = parallel.shard_games(total, r, world)
            seen += list(range(base, base + cnt))
        assert seen == list(range(total))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mx, sm = parallel.reduce_max_sum([10.0 + rank, 5.0 - rank], [100 * (rank + 1), 1])
        n = 3 + 2 * rank                       # ragged blocks: 3 and 5 rows
        base, _ = parallel.shard_games(8, rank, world)
        smp = dict(game=np.arange(base, base + n, dtype=np.int32), policy=np.full((n, 7), rank + 0.5, np.float32), state=np.full((n, 84), rank, np.int8))
        allg = parallel.gather_samples(smp)
        again = parallel.gather_samples(smp, reuse_buffers=True)       # (the flag only changes who owns the result on the NCCL path)
        assert all(np.array_equal(allg[k], again[k]) for k in allg)
        only0 = parallel.gather_samples(smp, dst=0)
        empty = parallel.gather_samples(dict(game=np.zeros(0 if rank == 0 else 2, np.int32)))
        q.put((rank, mx, sm, {k: v.tolist() for k, v in allg.items()}, None if only0 is None else len(only0["game"]), empty["game"].shape[0]))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_reduce_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mx, sm, allg, n0, nempty in res:
        assert mx == [11.0, 5.0] and sm == [300.0, 2.0]
        assert allg["game"] == [0, 1, 2, 4, 5, 6, 7, 8]          # rank 0 rows then rank 1 rows
        assert np.array(allg["policy"]).shape == (8, 7) and np.array(allg["policy"])[3:, 0].tolist() == [1.5] * 5
        assert np.array(allg["state"])[:3].max() == 0 and np.array(allg["state"])[3:].min() == 1
        assert n0 == (8 if rank == 0 else None)
        assert nempty == 2


def test_single_process_passthrough():
    smp = dict(game=np.arange(4, dtype=np.int32))
    assert parallel.gather_samples(smp) is smp
    assert parallel.reduce_max_sum([1.0], [2]) == ([1.0], [2.0])
