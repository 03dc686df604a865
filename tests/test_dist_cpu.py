"""Host-side multi-rank logic on CPU: contiguous sharding by global path index and the
[gradient | loss] sum-all-reduce, with world_size 2 over gloo.  No GPU, no kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deeppde_actorcritic_b200.solver import ActorCriticSolver
    s = object.__new__(ActorCriticSolver)          # host logic only: no engine, no device
    s.rank, s.world = rank, world
    lo, n = s._shard(B)
    # each rank contributes sum over its shard of f(global index): all-reduce must give the global sum
    idx = torch.arange(lo, lo + n, dtype=torch.float64)
    g = torch.stack([idx.sum(), (idx ** 2).sum()])
    loss = torch.tensor([float(n), 1.0], dtype=torch.float64)
    g2, loss2 = s._allreduce([g, loss])
    sh, lo2, Bg = s._shard_inputs((torch.arange(B), None, torch.arange(B)))
    q.put((rank, lo, n, g2.tolist(), loss2.tolist(), sh[0].tolist(), lo2, Bg))
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    B, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, lo, n, g, loss, sh0, lo2, Bg in out:
        covered += list(range(lo, lo + n))
        assert sh0 == list(range(lo, lo + n)) and lo2 == lo and Bg == B
        np.testing.assert_allclose(g, [sum(range(B)), sum(i * i for i in range(B))])
        np.testing.assert_allclose(loss, [B, world])
    assert covered == list(range(B))          # contiguous, disjoint, complete
