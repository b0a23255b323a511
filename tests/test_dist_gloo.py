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


# ---- DataParallelTrainer: flat buffers, hook-launched buckets, rank-invariant collectives ---------------------------
class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.emb = torch.nn.Parameter(torch.randn(50, 6))
        self.lin = torch.nn.Linear(7, 3)
        self.unused_on_rank1 = torch.nn.Parameter(torch.randn(5))


def _torch_adam_step(tr):
    """torch restatement of csrc/optim.cu (clip_grad_norm_ + Adam with L2 weight decay) for the CPU tests."""
    g = tr.flat_g
    if tr.grad_norm:
        coef = min(1.0, tr.grad_norm / (float(g.norm()) + 1e-6))
        g = g * coef
    g = g + tr.weight_decay * tr.flat_p
    b1, b2 = tr.betas
    tr.exp_avg.mul_(b1).add_(g, alpha=1 - b1)
    tr.exp_avg_sq.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** tr.step_count, 1 - b2 ** tr.step_count
    tr.flat_p.sub_((tr.lr / bc1) * tr.exp_avg / (tr.exp_avg_sq.sqrt() / bc2 ** 0.5 + tr.eps))


def _toy_loss(m, X, idx, use_extra):
    loss = (m.lin(X).pow(2).sum(1) + m.emb[idx].sum(1)).mean()
    if use_extra:
        loss = loss + m.unused_on_rank1.sum()
    return loss


def _trainer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    m = _Toy()
    tr = parallel.DataParallelTrainer(m, lr=1e-2, weight_decay=1e-3, grad_norm=0.5, bucket_bytes=256,
                                      optimizer_step=_torch_adam_step)
    assert len(tr.buckets) > 1
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        X, idx = torch.randn(12, 7, generator=g), torch.randint(0, 50, (12,), generator=g)
        lo, hi = parallel.shard_slice(12, rank, world)
        if step == 2 and rank == 1:
            # this rank's shard has no history at all: it must still take part in every collective
            def fn():
                raise ValueError('RGCNAggregator: every history in the batch is empty')
        else:
            # `unused_on_rank1` gets a gradient on rank 0 only: bucket sizes must not depend on that
            fn = lambda: _toy_loss(m, X[lo:hi], idx[lo:hi], use_extra=(rank == 0))     # noqa: E731
        tr.step(fn)
        tr.zero_grad()
    out[rank] = [p.detach().clone() for p in m.parameters()]
    dist.destroy_process_group()


def test_trainer_matches_single_process_global_batch():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_trainer_worker, args=(world, port, out), nprocs=world, join=True)
    # single-process reference: torch.optim.Adam on the averaged-over-ranks loss
    torch.manual_seed(0)
    m = _Toy()
    opt = torch.optim.Adam(m.parameters(), lr=1e-2, weight_decay=1e-3)
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        X, idx = torch.randn(12, 7, generator=g), torch.randint(0, 50, (12,), generator=g)
        losses = []
        for r in range(world):
            lo, hi = parallel.shard_slice(12, r, world)
            if step == 2 and r == 1:
                continue
            losses.append(_toy_loss(m, X[lo:hi], idx[lo:hi], use_extra=(r == 0)))
        (sum(losses) / world).backward()
        for p in m.parameters():          # parameters without a gradient take part with zeros (weight decay still applies)
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5)
        opt.step()
        opt.zero_grad()
    ref = [p.detach() for p in m.parameters()]
    for r in range(world):
        for a, b in zip(out[r], ref):
            assert torch.allclose(a, b, atol=1e-6), (r, (a - b).abs().max())
    for a, b in zip(out[0], out[1]):
        assert torch.equal(a, b)          # replicas stay bit-identical


def test_trainer_flat_views_alias_parameters():
    torch.manual_seed(0)
    m = _Toy()
    before = [p.detach().clone() for p in m.parameters()]
    tr = parallel.DataParallelTrainer(m, optimizer_step=_torch_adam_step)
    for p, b in zip(m.parameters(), before):
        assert torch.equal(p, b)
        assert p.data_ptr() >= tr.flat_p.data_ptr() and p.data_ptr() < tr.flat_p.data_ptr() + tr.flat_p.numel() * 4
        assert p.data_ptr() % 256 == tr.flat_p.data_ptr() % 256
        assert p.grad.data_ptr() >= tr.flat_g.data_ptr()
    tr.step(lambda: _toy_loss(m, torch.randn(4, 7), torch.randint(0, 50, (4,)), True))
    assert tr.flat_g.abs().sum() > 0
    tr.zero_grad()
    assert all(float(p.grad.abs().sum()) == 0 for p in m.parameters())
