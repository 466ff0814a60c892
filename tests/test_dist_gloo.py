"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: moment merge, logger reduction, gradient
all-reduce == single-process result on the concatenated batch, env sharding."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from egopose_b200 import dist_utils
    rng = np.random.RandomState(0)
    full = rng.randn(1000) * 3 + 1.5
    first, count = dist_utils.shard_envs(1000, rank, world)
    mine = full[first:first + count]
    stats = torch.tensor([len(mine), mine.mean(), ((mine - mine.mean()) ** 2).sum()], dtype=torch.float64)
    dist_utils.merge_moments_(stats)
    lg = torch.zeros(16, dtype=torch.float64)
    lg[0], lg[3], lg[4], lg[5], lg[11], lg[12] = count, mine.sum(), mine.min(), mine.max(), rank + 1, rank + 5
    dist_utils.reduce_logger_(lg, (4, 11), (5, 12))
    # gradient all-reduce: local mean-loss gradients scaled by the GLOBAL denominator sum to the global gradient
    W = torch.from_numpy(rng.randn(4, 3))
    X = torch.from_numpy(rng.randn(1000, 3))
    Y = torch.from_numpy(rng.randn(1000, 4))
    Xl, Yl = X[first:first + count], Y[first:first + count]
    g_local = 2 * (Xl @ W.t() - Yl).t() @ Xl / 1000.0
    dist_utils.allreduce_sum_(g_local)
    g_full = 2 * (X @ W.t() - Y).t() @ X / 1000.0
    q.put((rank, stats.tolist(), lg.tolist(), float((g_local - g_full).abs().max()), (first, count),
           [full.mean(), full.var(ddof=1), full.sum(), full.min(), full.max()]))
    dist.destroy_process_group()


def test_two_rank_host_logic():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    for rank, stats, lg, gerr, shard, ref in res:
        n, mean, m2 = stats
        assert n == 1000 and abs(mean - ref[0]) < 1e-12 and abs(m2 / (n - 1) - ref[1]) < 1e-11
        assert lg[0] == 1000 and abs(lg[3] - ref[2]) < 1e-9 and lg[4] == ref[3] and lg[5] == ref[4]
        assert lg[11] == 1 and lg[12] == 6
        assert gerr < 1e-12
    assert res[0][1] == res[1][1]                   # bit-identical merged moments on both ranks
    assert res[0][4] == (0, 500) and res[1][4] == (500, 500)


def test_shard_envs_uneven():
    from egopose_b200.dist_utils import shard_envs
    parts = [shard_envs(10, r, 4) for r in range(4)]
    assert parts == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert sum(c for _, c in parts) == 10
