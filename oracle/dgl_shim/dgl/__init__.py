"""TEST INFRASTRUCTURE ONLY -- a pure-torch stand-in for the DGL 0.4.x surface RE-Net uses.

The reference pins ``dgl-cuda10.1<0.5`` (reference README.md:38); DGL is not vendored in the
reference tree and is not installable here (no network).  This package restates the *published
DGL 0.4 semantics* of exactly the calls the reference makes on the hot path so that the reference's
own ``utils.py`` / ``RGCN.py`` / ``Aggregator.py`` / ``model.py`` can be imported UNCHANGED and used
as the parity oracle (see oracle/ref_loader.py).  Call sites it serves:

  utils.py:71-77     DGLGraph(), add_nodes, add_edges
  utils.py:90        in_degrees(range(n))
  utils.py:79-81     ndata.update / edata[...] =
  utils.py:121-124   g.subgraph(list) -> vertex-induced, relabelled in the given node order,
                     exposing ndata[dgl.NID] / edata[dgl.EID] (parent ids)
  utils.py:237       g.to(device)
  utils.py:238       dgl.batch(list)  (disjoint union, node ids offset, ndata/edata concatenated)
  RGCN.py:91         g.update_all(msg_udf, fn.sum(msg=, out=), apply_udf)
  Aggregator.py:59-61 dgl.max_nodes / dgl.mean_nodes (global aggregator only)

Semantics that matter for parity (DGL 0.4 documentation; "parity unpinned" by any reference test):
  * fn.sum sums messages over ALL in-edges of a node, parallel (multi-)edges included;
  * nodes with no in-edge receive zeros from the reduce;
  * the induced sub-graph keeps every edge (parallel edges too) whose two endpoints are selected;
  * a graph with zero edges skips message/reduce and only runs the apply UDF.
Nothing under oracle/ is imported by the product package (renet_b200/).
"""
import torch

from . import function  # noqa: F401

NID = '_ID'
EID = '_ID'


class _Frame(dict):
    """dict-like node/edge frame (update / pop / iteration over keys), like DGL's ndata/edata."""
    pass


class _EdgeBatch:
    def __init__(self, src, dst, data):
        self.src, self.dst, self.data = src, dst, data


class _NodeBatch:
    def __init__(self, data):
        self.data = data


class _LazyRows(dict):
    """frame view that gathers rows by an index on access (edges.src['h'] -> h[src_ids])."""

    def __init__(self, frame, index):
        super().__init__()
        self._frame, self._index = frame, index

    def __getitem__(self, k):
        return self._frame[k][self._index]


class DGLGraph:
    def __init__(self):
        self._n = 0
        self._src = torch.zeros(0, dtype=torch.long)
        self._dst = torch.zeros(0, dtype=torch.long)
        self.ndata = _Frame()
        self.edata = _Frame()
        self.batch_num_nodes = None

    # ---- construction -------------------------------------------------------------------
    def add_nodes(self, n):
        self._n += int(n)

    def add_edges(self, src, dst):
        src = torch.as_tensor(src, dtype=torch.long).view(-1)
        dst = torch.as_tensor(dst, dtype=torch.long).view(-1)
        self._src = torch.cat((self._src, src))
        self._dst = torch.cat((self._dst, dst))

    # ---- queries ------------------------------------------------------------------------
    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._src.numel())

    def edges(self):
        return self._src, self._dst

    def in_degrees(self, nodes=None):
        deg = torch.bincount(self._dst, minlength=self._n)
        if nodes is None:
            return deg
        return deg[torch.as_tensor(list(nodes), dtype=torch.long)]

    # ---- sub-graph / device -------------------------------------------------------------
    def subgraph(self, nodes):
        nodes = torch.as_tensor(list(nodes), dtype=torch.long)
        new_id = torch.full((self._n,), -1, dtype=torch.long)
        new_id[nodes] = torch.arange(nodes.numel())
        keep = (new_id[self._src] >= 0) & (new_id[self._dst] >= 0)
        eid = torch.nonzero(keep).view(-1)          # parent edge ids, ascending
        sg = DGLGraph()
        sg._n = int(nodes.numel())
        sg._src = new_id[self._src[eid]]
        sg._dst = new_id[self._dst[eid]]
        sg.ndata[NID] = nodes
        sg.edata[EID] = eid
        sg.parent_nid, sg.parent_eid = nodes, eid
        return sg

    def to(self, device):
        return self

    # ---- message passing ----------------------------------------------------------------
    def update_all(self, message_func, reduce_func, apply_node_func=None):
        if self.number_of_edges() > 0:
            src, dst = self._src.to(self._dev()), self._dst.to(self._dev())
            edges = _EdgeBatch(_LazyRows(self.ndata, src), _LazyRows(self.ndata, dst), self.edata)
            msgs = message_func(edges)
            reduce_func(self, msgs, dst)
        if apply_node_func is not None:
            self.ndata.update(apply_node_func(_NodeBatch(self.ndata)))

    def _dev(self):
        for v in self.ndata.values():
            return v.device
        return torch.device('cpu')


def batch(graph_list):
    bg = DGLGraph()
    off, srcs, dsts, sizes = 0, [], [], []
    for g in graph_list:
        srcs.append(g._src + off)
        dsts.append(g._dst + off)
        off += g._n
        sizes.append(g._n)
    bg._n = off
    bg._src = torch.cat(srcs) if srcs else torch.zeros(0, dtype=torch.long)
    bg._dst = torch.cat(dsts) if dsts else torch.zeros(0, dtype=torch.long)
    bg.batch_num_nodes = sizes
    if graph_list:
        for k in graph_list[0].ndata:
            if all(k in g.ndata for g in graph_list):
                bg.ndata[k] = torch.cat([g.ndata[k] for g in graph_list], dim=0)
        for k in graph_list[0].edata:
            if all(k in g.edata for g in graph_list):
                bg.edata[k] = torch.cat([g.edata[k] for g in graph_list], dim=0)
    return bg


def _segment_pool(g, feat, op):
    out, off = [], 0
    h = g.ndata[feat]
    sizes = g.batch_num_nodes if g.batch_num_nodes is not None else [g.number_of_nodes()]
    for n in sizes:
        seg = h[off:off + n]
        out.append(seg.max(dim=0)[0] if op == 'max' else seg.mean(dim=0))
        off += n
    return torch.stack(out)


def max_nodes(g, feat):
    return _segment_pool(g, feat, 'max')


def mean_nodes(g, feat):
    return _segment_pool(g, feat, 'mean')
