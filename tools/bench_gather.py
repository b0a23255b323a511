"""Micro-benchmark of the fused gather kernel variants on ICEWS18-shaped batches (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from renet_b200 import _lib, hoststore, synthetic  # noqa: E402

dev = torch.device('cuda:0')
L = _lib.lib()
P = _lib.ptr
tkg = synthetic.SyntheticTKG('icews18', seed=999, num_timestamps=240)
gs = hoststore.GraphStore(tkg.graph_dict)
hs = hoststore.HistoryStore(tkg.s_hist, tkg.s_hist_t, tkg.quads[:, 0], gs)
pool = []
for i in range(6):
    hb = hoststore.assemble_view(hs.select(tkg.batch_indices(i, 1024, tail_only=False)), dev, device_edges=False)
    g = hb.graph
    pool.append((hb, g, torch.randn(g.N, 200, device=dev), torch.empty(g.N, 200, device=dev)))
torch.manual_seed(0)
ent = torch.randn(tkg.num_e, 200, device=dev) * 0.1
W = torch.randn(512, 400, device=dev) * 0.1
stream = _lib.stream()
peak = 6562.6


def run(kind, variant, layer1):
    def call(hb, g, H, out):
        Hin, idx = (ent, g.node_ent) if layer1 else (H, None)
        if kind == 'comp':
            cptr, corder, slot, hot, n_hot = g.comp[False]
            rc = L.renet_rgcn_gather_comp(P(Hin), P(idx), P(W), P(g.row_ptr), P(g.col_src), P(g.col_type_s), P(g.norm),
                                          P(out), P(cptr), P(corder), P(slot), P(hot), n_hot, g.N, g.E, g.G, 200, 200,
                                          100, 512, 1, 1, stream)
        else:
            rc = L.renet_rgcn_gather(P(Hin), P(idx), P(W), P(g.row_ptr), P(g.col_src), P(g.col_type_s), P(g.norm), P(out),
                                     g.N, g.E, 200, 200, 100, 512, 1, 1, stream)
        _lib.check(rc, kind)
    L.renet_set_gather_variant(variant)
    for p in pool:
        call(*p)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 10
    for _ in range(reps):
        for p in pool:
            call(*p)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / (reps * len(pool)) * 1e3
    by = np.mean([g.E * 812 + g.N * 1604 + 512 * 1600 for _, g, _, _ in pool])
    print('%-5s variant %d layer%d: %7.1f us  %6.0f GB/s alg  %.1f%% of HBM peak' % (kind, variant, 1 if layer1 else 2, us,
                                                                                by / us / 1e3, by / us / 1e3 / peak * 100))


def check(v, layer1):
    """one clean call (self-loop input reset) -> output"""
    L.renet_set_gather_variant(v)
    outs = []
    for hb, g, H, out in pool[:2]:
        out.copy_(H * 0.5)
        Hin, idx = (ent, g.node_ent) if layer1 else (H, None)
        _lib.check(L.renet_rgcn_gather(P(Hin), P(idx), P(W), P(g.row_ptr), P(g.col_src), P(g.col_type_s), P(g.norm), P(out),
                                       g.N, g.E, 200, 200, 100, 512, 1, 1, stream), 'gather')
        torch.cuda.synchronize()
        outs.append(out.clone())
    return outs


variants = [int(x) for x in sys.argv[1:]] or [0, 1, 6]
for layer1 in (True, False):
    ref = check(variants[0], layer1)
    for v in variants:
        run('tile', v, layer1)
        got = check(v, layer1)
        print('      max |variant %d - variant %d| = %.3e (ref max %.3e)' % (v, variants[0], max(float((a - b).abs().max()) for a, b in zip(got, ref)), float(ref[0].abs().max())))
    # persistent kernel with the hot relation rows in shared memory (renet_set_hot_relations)
    for pairs in (24, 16):
        R = 256
        cnt = np.bincount(gs.type_s % R, minlength=R)
        top = np.argsort(-cnt, kind='stable')[:pairs]
        hot = np.ascontiguousarray(np.concatenate((top, top + R)).astype(np.int32))
        _lib.check(L.renet_set_hot_relations(hot.ctypes.data_as(_lib.ctypes.c_void_p), len(hot), 2 * R), 'hot')
        print('hot set: %d rows, edge share %.3f' % (len(hot), cnt[top].sum() / cnt.sum()))
        run('hot', 7, layer1)
        got = check(7, layer1)
        print('      max |hot - variant %d| = %.3e' % (variants[0], max(float((a - b).abs().max()) for a, b in zip(got, ref))))
        L.renet_set_hot_relations(None, 0, 0)
