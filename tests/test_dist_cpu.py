"""world_size-2 gloo test (CPU) of the data-parallel exchange step and batch sharding (mnexp_b200/dist.py)."""
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
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mnexp_b200.dist import exchange, shard_batch
    g = np.random.default_rng(0)
    glob = dict(user=g.integers(0, 10, 8).astype(np.int32), hist_doc=g.integers(0, 50, (8, 5)).astype(np.int32))
    mine = shard_batch(glob, rank, world)
    dense = torch.full((7,), float(rank + 1))
    ids = torch.as_tensor(mine['user'])
    rows = torch.full((4, 3), float(rank)) + torch.arange(4).float()[:, None]
    all_ids, all_rows = exchange(dense, ids, rows)
    q.put((rank, dense.numpy().copy(), all_ids.numpy().copy(), all_rows.numpy().copy(), mine['hist_doc'].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_and_sharding_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda r: r[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.random.default_rng(0)
    glob_user = g.integers(0, 10, 8).astype(np.int32)
    glob_hist = g.integers(0, 50, (8, 5)).astype(np.int32)
    for rank, dense, ids, rows, hist in res:
        assert np.all(dense == 3.0)                                   # 1 + 2: all-reduce(sum)
        assert np.array_equal(ids, glob_user)                         # rank-major gather reproduces the global batch
        assert np.array_equal(rows[:4, 0], np.arange(4)) and np.array_equal(rows[4:, 0], np.arange(4) + 1)
        assert np.array_equal(hist, glob_hist[rank * 4:(rank + 1) * 4])   # contiguous shards
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3])   # identical on every rank


def _worker_gather(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mnexp_b200.dist import gather_rows, shard_range, shard_users
    n_docs = 11                                                   # not a multiple of the world size
    lo, hi, per = shard_range(n_docs, rank, world)
    local = torch.zeros((per, 3))
    local[:hi - lo] = torch.arange(lo, hi).float()[:, None] + torch.tensor([0.0, 0.25, 0.5])
    table = gather_rows(local, n_docs)
    q.put((rank, (lo, hi, per), table.numpy().copy(), shard_users(7)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_inference_gather_world2():
    """C4 decomposed inference: documents sharded by id, one all-gather of the vectors, users sharded by id."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_gather, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda r: r[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(11, dtype=np.float32)[:, None] + np.array([0.0, 0.25, 0.5], dtype=np.float32)
    assert res[0][1] == (0, 6, 6) and res[1][1] == (6, 11, 6)
    for rank, _, table, users in res:
        assert table.shape == (11, 3) and np.array_equal(table, want)
    assert res[0][3] == (0, 4) and res[1][3] == (4, 7)
