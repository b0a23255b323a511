"""How fast can the host side alone deliver assembled batches? (GPU box)"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from renet_b200 import hoststore, synthetic
dev = torch.device('cuda:0')
tkg = synthetic.SyntheticTKG('icews18', seed=999, num_timestamps=240)
gs = hoststore.GraphStore(tkg.graph_dict)
hs_s = hoststore.HistoryStore(tkg.s_hist, tkg.s_hist_t, tkg.quads[:, 0], gs)
hs_o = hoststore.HistoryStore(tkg.o_hist, tkg.o_hist_t, tkg.quads[:, 2], gs)
sels = [tkg.batch_indices(i, 1024, tail_only=False) for i in range(8)]
buf = np.zeros(1 << 21, np.int32)
v = hs_s.select(sels[0])
for _ in range(3): hoststore.assemble_view_raw(v, buf)
t0 = time.perf_counter()
for i in range(20): hoststore.assemble_view_raw(hs_s.select(sels[i % 8]), buf)
print('single-thread assemble_view_raw: %.2f ms' % ((time.perf_counter() - t0) / 20 * 1e3))
for workers, depth, inner in ((4, 2, 8), (4, 2, 2), (8, 4, 1), (8, 4, 2), (12, 6, 1), (16, 8, 1)):
    groups = [(hs_s.select(sels[i % 8]), hs_o.select(sels[i % 8])) for i in range(44)]
    n = 0
    for hbs in hoststore.prefetch(iter(groups[:4]), dev, depth=depth, workers=workers, inner_threads=inner): pass
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for hbs in hoststore.prefetch(iter(groups[4:]), dev, depth=depth, workers=workers, inner_threads=inner): n += 1
    torch.cuda.synchronize()
    print('prefetch workers=%d depth=%d inner=%d: %.2f ms per step (2 directions, incl. H2D)' % (workers, depth, inner, (time.perf_counter() - t0) / n * 1e3))
