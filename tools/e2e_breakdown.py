"""Where does an e2e step go?  (GPU box)  python tools/e2e_breakdown.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from renet_b200 import hoststore, synthetic
from renet_b200.model import RENet

dev = torch.device('cuda:0')
tkg = synthetic.SyntheticTKG('icews18', seed=999, num_timestamps=240)
gs = hoststore.GraphStore(tkg.graph_dict)
hs_s = hoststore.HistoryStore(tkg.s_hist, tkg.s_hist_t, tkg.quads[:, 0], gs)
hs_o = hoststore.HistoryStore(tkg.o_hist, tkg.o_hist_t, tkg.quads[:, 2], gs)
m = RENet(tkg.num_e, 200, tkg.num_r, dropout=0).to(dev).eval()
m.global_emb = {t: v.to(dev) for t, v in tkg.global_emb.items()}
sels = [tkg.batch_indices(i, 1024, tail_only=False) for i in range(8)]
qp = [torch.from_numpy(tkg.quads[s]).pin_memory() for s in sels]
ring = [torch.empty(2048, 400).pin_memory() for _ in range(2)]
N = 24


def loop(n, encode=True, workers=4, depth=2, dev_edges=True):
    idx = [i % len(sels) for i in range(n)]
    groups = ((hs_s.select(sels[j]), hs_o.select(sels[j])) for j in idx)
    t = dict(fetch=0.0, enc=0.0, d2h=0.0)
    prev = None
    it = hoststore.prefetch(groups, dev, depth=depth, workers=workers, device_edges=dev_edges)
    for i, j in enumerate(idx):
        t0 = time.perf_counter()
        hbs = next(it)
        t1 = time.perf_counter()
        if encode:
            batch = qp[j].to(dev, non_blocking=True)
            outs = []
            with torch.no_grad():
                for subj in (True, False):
                    s, r, o, s_h, s_q, _ = m.encode(batch, hbs[0], hbs[1], gs, subject=subj)
                    outs.append(torch.cat((s_h, s_q), 1))
            res = torch.cat(outs)
            t2 = time.perf_counter()
            ring[i & 1][:res.shape[0]].copy_(res, non_blocking=True)
            ev = torch.cuda.Event(); ev.record()
            if prev is not None:
                prev.synchronize()
            prev = ev
            t3 = time.perf_counter()
        else:
            t2 = t3 = t1
        t['fetch'] += t1 - t0; t['enc'] += t2 - t1; t['d2h'] += t3 - t2
    for _ in it:
        pass
    torch.cuda.synchronize()
    return {k: v / n * 1e3 for k, v in t.items()}


for dev_edges in (True, False):
    for workers, depth in ((4, 2), (8, 4)):
        loop(6, dev_edges=dev_edges, workers=workers, depth=depth)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        br = loop(N, dev_edges=dev_edges, workers=workers, depth=depth)
        full = (time.perf_counter() - t0) / N * 1e3
        torch.cuda.synchronize(); t0 = time.perf_counter()
        loop(N, encode=False, dev_edges=dev_edges, workers=workers, depth=depth)
        deliver = (time.perf_counter() - t0) / N * 1e3
        print('device_edges=%s workers=%d depth=%d: full %.2f ms/step (fetch %.2f enc %.2f d2h+wait %.2f) ; delivery only %.2f ms/step'
              % (dev_edges, workers, depth, full, br['fetch'], br['enc'], br['d2h'], deliver))
# GPU-only
for dev_edges in (True, False):
    hbs_all = [(hoststore.assemble_view(hs_s.select(sels[j]), dev, device_edges=dev_edges),
                hoststore.assemble_view(hs_o.select(sels[j]), dev, device_edges=dev_edges)) for j in range(4)]
    batch = qp[0].to(dev)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        t0 = time.perf_counter(); a.record()
        with torch.no_grad():
            for k in range(12):
                for subj in (True, False):
                    m.encode(batch, hbs_all[k % 4][0], hbs_all[k % 4][1], gs, subject=subj)
        b.record(); torch.cuda.synchronize()
    print('encode x2 on resident batches (device_edges=%s): GPU %.2f ms, wall %.2f ms per step' % (dev_edges, a.elapsed_time(b) / 12, (time.perf_counter() - t0) / 12 * 1e3))
# single-thread host costs
buf = np.zeros(1 << 21, np.int32)
for name, fn in (('plan', hoststore.plan_view_raw), ('assemble', hoststore.assemble_view_raw)):
    for _ in range(3): fn(hs_s.select(sels[0]), buf)
    t0 = time.perf_counter()
    for i in range(16): fn(hs_s.select(sels[i % 8]), buf)
    print('%s_view_raw single call: %.2f ms' % (name, (time.perf_counter() - t0) / 16 * 1e3))
# upload cost on the consumer thread
for dev_edges in (True, False):
    views = [hs_s.select(sels[j]) for j in range(8)]
    holders = [[torch.empty(1 << 21, dtype=torch.int32).pin_memory()] for _ in range(8)]
    rs = [hoststore._stage(v, h, True, dev_edges) for v, h in zip(views, holders)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for rep in range(4):
        for v, h, r in zip(views, holders, rs):
            hoststore._upload(v, h[0], r, dev)
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print('_upload device_edges=%s: %.3f ms host per batch (%.3f incl. GPU drain)' % (dev_edges, (t1 - t0) / 32 * 1e3, (time.perf_counter() - t0) / 32 * 1e3))

if '--trace' in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    loop(6)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        loop(12)
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in evs)
    span = iv[-1][1] - iv[0][0]
    busy, cur_s, cur_e = 0, iv[0][0], iv[0][1]
    gaps = []
    for s, e, n in iv[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, n))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    print('GPU timeline over 12 steps: span %.2f ms, busy %.2f ms (%.0f%%), %d gaps' % (span / 1e3, busy / 1e3, 100.0 * busy / span, len(gaps)))
    gaps.sort(reverse=True)
    print('largest gaps (us, next kernel):', [(round(g, 1), n[:40]) for g, n in gaps[:12]])
    agg = {}
    for s, e, n in iv:
        agg[n[:60]] = agg.get(n[:60], 0) + (e - s)
    for n, t in sorted(agg.items(), key=lambda x: -x[1])[:14]:
        print('  %8.1f us/step  %s' % (t / 12, n))
