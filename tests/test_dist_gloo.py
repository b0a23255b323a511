"""-m "not gpu": the N>1 host logic on CPU with gloo, world_size 2 (sharding + weighted gradient all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from renet_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(7, 3)
    emb = torch.nn.Parameter(torch.randn(50, 6))
    params = list(lin.parameters()) + [emb]
    # global batch of 11 samples, ragged shards (6 + 5); loss = mean over LOCAL samples like nn.CrossEntropyLoss
    g = torch.Generator().manual_seed(1)
    X, idx = torch.randn(11, 7, generator=g), torch.randint(0, 50, (11,), generator=g)
    lo, hi = parallel.shard_slice(11, rank, world)
    loss = (lin(X[lo:hi]).pow(2).sum(1) + emb[idx[lo:hi]].sum(1)).mean()
    loss.backward()
    parallel.allreduce_gradients(params, local_weight=hi - lo, bucket_bytes=64)   # tiny buckets: several collectives
    out[rank] = [p.grad.clone() for p in params]
    dist.destroy_process_group()


def test_weighted_gradient_allreduce_equals_global_batch_gradient():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    lin = torch.nn.Linear(7, 3)
    emb = torch.nn.Parameter(torch.randn(50, 6))
    g = torch.Generator().manual_seed(1)
    X, idx = torch.randn(11, 7, generator=g), torch.randint(0, 50, (11,), generator=g)
    (lin(X).pow(2).sum(1) + emb[idx].sum(1)).mean().backward()
    ref = [p.grad for p in list(lin.parameters()) + [emb]]
    for r in range(world):
        for a, b in zip(out[r], ref):
            assert torch.allclose(a, b, atol=1e-6), r


def test_shard_slices_cover_the_batch():
    for n in (8192, 1000, 7):
        for world in (1, 2, 4, 8):
            cuts = [parallel.shard_slice(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1
    b = np.arange(10)
    sh = ([[i] for i in range(10)], [[i] for i in range(10)])
    bb, s, o, n = parallel.shard_batch(b, sh, sh, 1, 3)
    assert list(bb) == [4, 5, 6] and s[0] == [[4], [5], [6]] and n == 3
