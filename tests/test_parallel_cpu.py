"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (mdctgan_b200/parallel.py): bucket layout, the
single sum all-reduce + 1/world scale = gradient of the mean loss over the global batch, replica sync, segment sharding.
(The reference has no multi-GPU semantics; the contract checked here is "same math as one process on the
global batch".)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mdctgan_b200 import parallel as P

    torch.manual_seed(100)                      # the same "model" on every rank
    shapes = [(8, 3, 3, 3), (8,), (5, 8, 1, 1), (5,)]
    offs, total = P.flat_layout([int(np.prod(s)) for s in shapes])
    w = torch.randn(total, dtype=torch.float64)
    # a different data shard per rank: per-rank loss = mean over the local batch of (w . x)^2 -> gradient
    g_global = torch.Generator().manual_seed(7)
    x_all = torch.randn(world * 4, total, dtype=torch.float64, generator=g_global)
    mine = P.shard_range(world * 4, world, rank)
    x = x_all[mine.start:mine.stop]
    grad_local = (2 * (x @ w)[:, None] * x).mean(0)
    ex = P.GradExchange()
    flat = grad_local.clone()
    ex(flat)
    grad_dp = flat * ex.grad_scale
    grad_ref = (2 * (x_all @ w)[:, None] * x_all).mean(0)      # one process on the global batch
    ok_grad = torch.allclose(grad_dp, grad_ref, rtol=1e-12, atol=1e-14)
    # replica sync + broadcast
    p = w.clone() + (0.0 if rank == 0 else 1.0)
    in_sync_before = P.check_replicas_in_sync(p)
    P.broadcast_flat([p])
    in_sync_after = P.check_replicas_in_sync(p)
    out[rank] = dict(ok_grad=bool(ok_grad), calls=ex.calls, scale=ex.grad_scale, before=in_sync_before, after=in_sync_after,
                     shard=(mine.start, mine.stop))
    dist.destroy_process_group()


def test_flat_layout_and_shards():
    sys.path.insert(0, ROOT)
    from mdctgan_b200 import parallel as P

    offs, total = P.flat_layout([5, 8, 1, 16])
    assert offs == [0, 8, 16, 20] and total == 36
    for n, w in ((89, 8), (7, 8), (32, 4), (0, 2)):
        covered = [i for r in range(w) for i in P.shard_range(n, w, r)]
        assert covered == list(range(n))
        sizes = [len(P.shard_range(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_single_allreduce_equals_global_batch_gradient_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        o = out[r]
        assert o["ok_grad"] and o["calls"] == 1 and o["scale"] == 0.5
        assert o["before"] is False and o["after"] is True
    assert out[0]["shard"] == (0, 4) and out[1]["shard"] == (4, 8)
