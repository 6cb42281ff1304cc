"""Host-side multi-GPU logic on CPU: contiguous sharding and the best-cost gather with
world_size 2 over gloo (the GPU path uses the same code over NCCL)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from or_cdchomp_b200 import sharding


def test_shard_bounds_cover_everything():
    for n in (1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_local_best_skips_failed_and_nan():
    c = np.array([5.0, 1.0, np.nan, 0.5, 0.5])
    s = np.array([0, -5, 0, 0, 0])
    assert sharding.local_best(c, s, 100) == (0.5, 103)
    assert sharding.local_best(np.array([1.0]), np.array([-5]), 0) == (float("inf"), -1)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    total, P, n = 37, 6, 3
    costs = rng.uniform(1, 2, size=total)
    costs[29] = 0.25
    costs[30] = 0.25  # tie: the lower global id must win
    trajs = rng.normal(size=(total, P, n))
    status = np.zeros(total, dtype=np.int32)
    status[3] = -5
    lo, hi = sharding.shard_bounds(total, rank, world)
    c, gid = sharding.local_best(costs[lo:hi], status[lo:hi], lo)
    t = torch.from_numpy(trajs[gid].copy()) if gid >= 0 else torch.zeros(P, n, dtype=torch.float64)
    wc, wid, wt = sharding.gather_best(c, gid, t, P, n, torch.device("cpu"))
    q.put((rank, wc, wid, np.abs(wt.numpy() - trajs[29]).max()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_best_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, wc, wid, err in res:
        assert wc == 0.25 and wid == 29 and err == 0.0
