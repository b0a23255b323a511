"""Per-kernel durations of one fused GRU forward (run under ncu on the GPU box)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from renet_b200 import hoststore, synthetic, utils
from renet_b200.gru import fused_gru
from renet_b200.model import RENet
dev = torch.device('cuda:0')
tkg = synthetic.SyntheticTKG('icews18', seed=999, num_timestamps=240)
gs = hoststore.GraphStore(tkg.graph_dict)
hs = hoststore.HistoryStore(tkg.s_hist, tkg.s_hist_t, tkg.quads[:, 0], gs)
m = RENet(tkg.num_e, 200, tkg.num_r).to(dev).eval()
m.global_emb = {t: v.to(dev) for t, v in tkg.global_emb.items()}
sel = tkg.batch_indices(0, 1024, tail_only=False)
hb = hoststore.assemble_view(hs.select(sel), dev)
H2 = torch.randn(hb.graph.N, 200, device=dev)
glob = utils.global_rows(m.global_emb, hb.times, 200, dev)
q = tkg.quads[sel]
s_tem = torch.from_numpy(q[:, 0][hb.s_idx]).to(dev).to(torch.int32)[:hb.num_seq].contiguous()
r_tem = torch.from_numpy(q[:, 1][hb.s_idx]).to(dev).to(torch.int32)[:hb.num_seq].contiguous()
with torch.no_grad():
    for i in range(3):
        if i == 2:
            torch.cuda.synchronize(); torch.cuda.nvtx.range_push('gru')
        fused_gru(H2, m.ent_embeds, m.rel_embeds[:tkg.num_r], glob, hb, s_tem, r_tem, m.encoder, m.encoder_r)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fused_gru(H2, m.ent_embeds, m.rel_embeds[:tkg.num_r], glob, hb, s_tem, r_tem, m.encoder, m.encoder_r)
    b.record(); torch.cuda.synchronize()
    print('fused_gru %.1f us' % (a.elapsed_time(b) / 10 * 1e3), 'S', hb.S, 'Q', hb.num_seq)
