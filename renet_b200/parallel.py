"""Data parallelism for the hot path: one process per GPU, every rank runs the
whole RGCN-aggregate + GRU path on its own slice of the global batch (its own batched history graph --
components may repeat across ranks, which is cheaper than exchanging them), parameters replicated, and ONE
flattened gradient all-reduce per step (NCCL over NVLink on the GPUs; gloo on CPU for the tests), followed by
clip-grad-norm on the reduced gradient (reference train.py:140) and the optimiser step.

The reference itself is single-GPU (train.py:33); nothing here has a reference counterpart.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_slice(n, rank, world):
    """[lo, hi) of rank's contiguous slice of n items (global batch 8192 -> 8 x 1024)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch, s_hist, o_hist, rank, world):
    """Slice (triplets, (s_hist, s_hist_t), (o_hist, o_hist_t)) of a GLOBAL batch for this rank."""
    lo, hi = shard_slice(len(batch), rank, world)
    cut = lambda pair: (pair[0][lo:hi], pair[1][lo:hi])
    return batch[lo:hi], cut(s_hist), cut(o_hist), (hi - lo)


def allreduce_gradients(params, local_weight=1.0, bucket_bytes=64 << 20, group=None):
    """Average gradients over ranks with a weight per rank (ragged last batch: weight = local sample count;
    nn.CrossEntropyLoss is a mean over LOCAL samples, model.py:57).  Gradients are flattened into buckets of
    ``bucket_bytes`` so the collective count stays small (80.9 MB of fp32 parameters -> 2 buckets) and
    bucket i+1 is packed while bucket i is in flight (async_op)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    dev = params[0].grad.device
    w = torch.tensor([float(local_weight)], device=dev, dtype=torch.float32)
    dist.all_reduce(w, group=group)
    scale = float(local_weight) / float(w.item())
    buckets, cur, cur_bytes = [], [], 0
    for p in params:
        nb = p.grad.numel() * p.grad.element_size()
        if cur and cur_bytes + nb > bucket_bytes:
            buckets.append(cur); cur, cur_bytes = [], 0
        cur.append(p); cur_bytes += nb
    if cur:
        buckets.append(cur)
    pending = []
    for b in buckets:
        flat = torch.cat([p.grad.reshape(-1) for p in b]).mul_(scale)
        pending.append((b, flat, dist.all_reduce(flat, group=group, async_op=True)))
    for b, flat, work in pending:
        work.wait()
        off = 0
        for p in b:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n


def train_step(model, optimizer, batch, s_hist, o_hist, graph_dict, grad_norm=1.0, local_weight=None):
    """One reference training step (train.py:136-143) on this rank's shard + the gradient all-reduce."""
    loss_s = model(batch, s_hist, o_hist, graph_dict, subject=True)
    loss_o = model(batch, s_hist, o_hist, graph_dict, subject=False)
    loss = loss_s + loss_o
    loss.backward()
    allreduce_gradients(list(model.parameters()), local_weight if local_weight is not None else len(batch))
    torch.nn.utils.clip_grad_norm_(model.parameters(), grad_norm)
    optimizer.step()
    optimizer.zero_grad()
    return loss.detach()
