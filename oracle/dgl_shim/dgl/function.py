"""TEST INFRASTRUCTURE ONLY -- stand-in for ``dgl.function`` (reference RGCN.py:3,91 uses fn.sum)."""
import torch


class _Sum:
    """fn.sum(msg=..., out=...): out[v] = sum of msg over all in-edges of v (zeros if none)."""

    def __init__(self, msg, out):
        self.msg, self.out = msg, out

    def __call__(self, g, msgs, dst):
        m = msgs[self.msg]
        acc = torch.zeros((g.number_of_nodes(),) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
        acc = acc.index_add(0, dst, m)
        g.ndata[self.out] = acc


def sum(msg, out):  # noqa: A001 - mirrors dgl.function.sum
    return _Sum(msg, out)
