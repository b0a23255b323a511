"""debug: per-call comparison of aggregator.encode (CUDA) against the oracle encode during the eval flow."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import eval_setup, eval_flow, rel_err
import test_inference_host as T

ctx = eval_setup('cuda:0')
m = ctx['model']
oracle = T._oracle_encode(ctx)
real = m.aggregator.encode
n = [0, 0]

def both(hist, s, r, ent, rel, gd, ge, reverse, enc, encr):
    a = real(hist, s, r, ent, rel, gd, ge, reverse, enc, encr)
    b = oracle(hist, s, r, ent, rel, gd, ge, reverse, enc, encr)
    e1, e2 = rel_err(a[0].cpu().numpy(), b[0].numpy()), rel_err(a[1].cpu().numpy(), b[1].numpy())
    n[0] += 1
    if max(e1, e2) > 1e-4:
        n[1] += 1
        if n[1] <= 6:
            print('MISMATCH call', n[0], 'Q', len(hist[0]), 'lens', [len(x) for x in hist[0]][:4], 't', hist[1][0], 's', s.view(-1)[:3].tolist(),
                  'r', r.view(-1)[:3].tolist(), 'reverse', reverse, 'err', e1, e2)
            print('   hist sizes', [np.asarray(x).shape for x in hist[0][0]])
    return a
m.aggregator.encode = both
res = eval_flow(ctx, 'cuda:0')
print('calls', n)
