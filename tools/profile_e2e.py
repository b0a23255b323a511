"""Host-side breakdown of one e2e step (run on the GPU box): python tools/profile_e2e.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from renet_b200 import hoststore, synthetic  # noqa: E402
from renet_b200.model import RENet  # noqa: E402

dev = torch.device('cuda:0')
tkg = synthetic.SyntheticTKG('icews18', seed=999, num_timestamps=240)
gs = hoststore.GraphStore(tkg.graph_dict)
hs_s = hoststore.HistoryStore(tkg.s_hist, tkg.s_hist_t, tkg.quads[:, 0], gs)
hs_o = hoststore.HistoryStore(tkg.o_hist, tkg.o_hist_t, tkg.quads[:, 2], gs)
m = RENet(tkg.num_e, 200, tkg.num_r, dropout=0).to(dev).eval()
m.global_emb = {t: v.to(dev) for t, v in tkg.global_emb.items()}
sels = [tkg.batch_indices(i, 1024, tail_only=False) for i in range(6)]


def run(n, first):
    idx = [(first + i) % len(sels) for i in range(n)]
    groups = ((hs_s.select(sels[j]), hs_o.select(sels[j])) for j in idx)
    for j, hbs in zip(idx, hoststore.prefetch(groups, dev, depth=2, workers=4)):
        batch = torch.from_numpy(tkg.quads[sels[j]]).pin_memory().to(dev, non_blocking=True)
        outs = []
        with torch.no_grad():
            for subj in (True, False):
                s, r, o, s_h, s_q, _ = m.encode(batch, hbs[0], hbs[1], gs, subject=subj)
                outs.append(torch.cat((s_h, s_q), 1))
        torch.cat(outs).cpu()


run(4, 0)
torch.cuda.synchronize()
t0 = time.perf_counter()
run(20, 0)
torch.cuda.synchronize()
print('e2e step (prefetch) %.2f ms' % ((time.perf_counter() - t0) / 20 * 1e3))
# GPU-only time of the same work
hbs_all = [(hoststore.assemble_view(hs_s.select(sels[j]), dev), hoststore.assemble_view(hs_o.select(sels[j]), dev)) for j in range(3)]
batch = torch.from_numpy(tkg.quads[sels[0]]).to(dev)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
with torch.no_grad():
    for k in range(10):
        for subj in (True, False):
            m.encode(batch, hbs_all[k % 3][0], hbs_all[k % 3][1], gs, subject=subj)
b.record(); torch.cuda.synchronize()
print('encode x2 without batching/copies: GPU %.2f ms, wall %.2f ms per step' % (a.elapsed_time(b) / 10, (time.perf_counter() - t0) / 10 * 1e3))
pr = cProfile.Profile()
pr.enable()
run(10, 0)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(22)
