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


def step(i):
    sel = sels[i % len(sels)]
    batch = torch.from_numpy(tkg.quads[sel]).pin_memory().to(dev, non_blocking=True)
    outs = []
    with torch.no_grad():
        for subj in (True, False):
            s, r, o, s_h, s_q, _ = m.encode(batch, hs_s.select(sel), hs_o.select(sel), gs, subject=subj)
            outs.append(torch.cat((s_h, s_q), 1))
    return torch.cat(outs).cpu()


for i in range(3):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    step(i)
torch.cuda.synchronize()
print('e2e step %.2f ms' % ((time.perf_counter() - t0) / 10 * 1e3))
pr = cProfile.Profile()
pr.enable()
for i in range(10):
    step(i)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(28)
