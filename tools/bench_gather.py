"""Micro-benchmark + cross-check of the fused gather kernels on dataset-shaped batches (run on the GPU box).

    RENET_GATHER_KERNEL=tile   python tools/bench_gather.py [icews18|gdelt|icews14] [timestamps]
    RENET_GATHER_KERNEL=stream python tools/bench_gather.py ...

The kernel choice is read once per process (environment), so an A/B is two runs; each run also writes a checksum of the
layer outputs so the two kernels can be compared (they agree to fp32 summation order, not bit for bit)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from renet_b200 import _lib, hoststore, synthetic  # noqa: E402

preset = sys.argv[1] if len(sys.argv) > 1 else 'icews18'
T = int(sys.argv[2]) if len(sys.argv) > 2 else {'icews18': 240, 'gdelt': 2138, 'icews14': 181}[preset]
dev = torch.device('cuda:0')
L = _lib.lib()
P = _lib.ptr
tkg = synthetic.SyntheticTKG(preset, seed=999, num_timestamps=T)
R2 = 2 * tkg.num_r
gs = hoststore.GraphStore(tkg.graph_dict)
hs = hoststore.HistoryStore(tkg.s_hist, tkg.s_hist_t, tkg.quads[:, 0], gs)
torch.manual_seed(123)
pool = []
for i in range(6):
    hb = hoststore.assemble_view(hs.select(tkg.batch_indices(i, 1024, tail_only=False)), dev, device_edges=False)
    g = hb.graph
    pool.append((hb, g, torch.randn(g.N, 200, device=dev), torch.empty(g.N, 200, device=dev)))
torch.manual_seed(0)
ent = torch.randn(tkg.num_e, 200, device=dev) * 0.1
W = torch.randn(R2, 400, device=dev) * 0.1
stream = _lib.stream()
peak = 6562.6
kernel = os.environ.get('RENET_GATHER_KERNEL', 'auto')


# RENET_HOT_LIST=1: pass the dataset's relation ranking (all of graph_dict's type_s columns) instead of per-CTA ranking
hot = None
if os.environ.get('RENET_HOT_LIST', '0') == '1':
    freq = np.zeros(R2, dtype=np.int64)
    for gg in tkg.graph_dict.values():
        freq += np.bincount(np.asarray(gg.type_s, dtype=np.int64), minlength=R2)
    hot = torch.from_numpy(np.argsort(-freq, kind='stable')[:128].astype(np.int32)).to(dev)


def call(hb, g, H, out, layer1):
    Hin, idx = (ent, g.node_ent) if layer1 else (H, None)
    if hot is None:
        rc = L.renet_rgcn_gather(P(Hin), P(idx), P(W), P(g.row_ptr), P(g.col_src), P(g.col_type_s), P(g.norm), P(out),
                                 g.N, g.E, 200, 200, 100, R2, 1, 1, stream)
    else:
        rc = L.renet_rgcn_gather_hot(P(Hin), P(idx), P(W), P(g.row_ptr), P(g.col_src), P(g.col_type_s), P(g.norm), P(out),
                                     g.N, g.E, 200, 200, 100, R2, 1, 1, P(hot), hot.numel(), stream)
    _lib.check(rc, 'renet_rgcn_gather')


for layer1 in (True, False):
    for p in pool:
        call(*p, layer1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 20
    for _ in range(reps):
        for p in pool:
            call(*p, layer1)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / (reps * len(pool)) * 1e3
    by = np.mean([g.E * 812 + g.N * 1604 + R2 * 1600 for _, g, _, _ in pool])
    print('%s %-6s layer%d: N %d E %d  %7.1f us  %6.0f GB/s alg  %.1f%% of HBM peak' % (
        preset, kernel, 1 if layer1 else 2, pool[0][1].N, pool[0][1].E, us, by / us / 1e3, by / us / 1e3 / peak * 100))

# one clean call per layer kind -> checksum + file for the A/B comparison
hb, g, H, out = pool[0]
res = {}
for layer1 in (True, False):
    out.copy_(torch.arange(g.N * 200, device=dev, dtype=torch.float32).view(g.N, 200).remainder(7.0) * 0.01)
    call(hb, g, H, out, layer1)
    torch.cuda.synchronize()
    res['layer%d' % (1 if layer1 else 2)] = out.cpu().numpy().copy()
    first = out.clone()
    out.copy_(torch.arange(g.N * 200, device=dev, dtype=torch.float32).view(g.N, 200).remainder(7.0) * 0.01)
    call(hb, g, H, out, layer1)
    print('  reproducible bit for bit:', bool(torch.equal(first, out)), ' checksum %.6f' % float(out.double().sum()))
os.makedirs('gpurun_out', exist_ok=True)
path = '/tmp/gather_%s_%s.npz' % (preset, kernel)
np.savez(path, **res)
other = '/tmp/gather_%s_%s.npz' % (preset, 'tile' if kernel != 'tile' else 'stream')
if os.path.exists(other):
    o = np.load(other)
    for k in res:
        d = np.abs(res[k] - o[k]).max() / np.abs(o[k]).max()
        print('  vs %s %s: max rel diff %.3g' % (other, k, d))
