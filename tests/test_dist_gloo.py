"""CPU, world_size=2, gloo: the host-side sharding logic of the multi-GPU path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_tf2_b200 import dist as nd
    n = 1001
    a, b = nd.shard_range(n, rank, world)
    rows = torch.arange(a, b, dtype=torch.float32)[:, None].repeat(1, 5)     # [n_local, 5] "image rows"
    full = nd.gather_rows(rows, n)
    ok1 = torch.equal(full[:, 0], torch.arange(n, dtype=torch.float32))
    on0 = nd.gather_rows(rows, n, dst=0)
    ok2 = (on0 is not None) == (rank == 0)
    # data-parallel gradient: per-rank partial sums of a global-mean loss all-reduce to the global gradient
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.normal(size=(64, 8)).astype(np.float32))
    w = torch.zeros(8, requires_grad=True)
    lo, hi = nd.shard_range(64, rank, world)
    loss = ((x[lo:hi] @ w - 1.0) ** 2).sum() / 64.0                         # scaled by 1/B_global
    g = torch.autograd.grad(loss, w)[0]
    nd.allreduce_flat(g)
    gl = torch.autograd.grad(((x @ w - 1.0) ** 2).mean(), w)[0]
    ok3 = torch.allclose(g, gl, atol=1e-6)
    ok4 = nd.max_over_ranks(rank + 1.5, "cpu") == world + 0.5
    out[rank] = bool(ok1 and ok2 and ok3 and ok4)
    dist.destroy_process_group()


def test_shard_range_partitions():
    from nerf_tf2_b200 import dist as nd
    for n in (0, 1, 7, 640000, 4096):
        for world in (1, 2, 3, 8):
            spans = [nd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert sorted(sum((nd.shard_views(200, r, 8) for r in range(8)), [])) == list(range(200))


def test_two_rank_gloo_gather_and_gradient_allreduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
