"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: sharding, bucketed gradient
all-reduce, and the fusion-training reduction (sum of per-rank gradient accumulators == single-process
gradient over the concatenated points, checked with the numpy oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from multiplanarunet_b200 import distributed as D
    from oracle import fusion
    D.init_from_env(backend="gloo")
    out = {}
    out["shard"] = D.shard(list(range(7)))
    # bucketed all-reduce of a flat "gradient" buffer
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    D.all_reduce_flat(g, bucket_elems=300)
    out["g_ok"] = bool(torch.equal(g, torch.arange(1000, dtype=torch.float32) * 3))
    # identical parameters after broadcast
    p = torch.full((10,), float(rank))
    D.broadcast_flat(p, src=0)
    out["bcast_ok"] = bool((p == 0).all())
    # fusion training: each rank owns half of the points; reduce [dW | db | loss_sum] accumulators
    rng = np.random.RandomState(5)
    X = rng.rand(400, 6, 5).astype(np.float32)
    X /= X.sum(-1, keepdims=True)
    y = rng.randint(0, 5, 400)
    W = rng.uniform(0.5, 1.5, (6, 5))
    b = 0.1 * rng.randn(5)
    sl = slice(rank * 200, (rank + 1) * 200)
    loss, dW, db = fusion.gdl_loss_and_grads(X[sl], y[sl], W, b, reg=0.0)
    acc = torch.tensor(np.concatenate([dW.ravel() * 200, db * 200, [loss * 200]]))
    dist.all_reduce(acc)
    lf, dWf, dbf = fusion.gdl_loss_and_grads(X, y, W, b, reg=0.0)
    got = acc.numpy() / 400
    out["fusion_ok"] = bool(np.allclose(got[:30], dWf.ravel(), atol=1e-12) and np.allclose(got[30:35], dbf, atol=1e-12)
                            and abs(got[35] - lf) < 1e-12)
    q.put((rank, out))
    dist.destroy_process_group()


def test_gloo_world2_data_parallel_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["shard"] == [0, 2, 4, 6] and res[1]["shard"] == [1, 3, 5]
    for r in (0, 1):
        assert res[r]["g_ok"] and res[r]["bcast_ok"] and res[r]["fusion_ok"], res[r]
