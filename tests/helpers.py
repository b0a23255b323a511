"""Shared test helpers (CPU side).  The oracle is the checker; nothing here is product code."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def layer_case(blob, name):
    pre = name + '/'
    return {k[len(pre):]: blob[k] for k in blob.files if k.startswith(pre)}


def rel_err(a, b):
    """max-abs-diff / max-abs-ref (SURVEY.md section 8(d) parity definition)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(np.abs(b).max() if b.size else 0.0, 1e-30)
    return float(np.abs(a - b).max() / denom) if b.size else 0.0


def t(x, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(dtype) if dtype is not None else x


# ---- test-time path (model.py:107-446) ---------------------------------------------------------------------------------
def eval_setup(device):
    """The tiny TKG of tests/golden/renet_eval_tiny.npz on the product RENet (parameters, graphs, histories, state)."""
    from oracle import restate
    from oracle.gen_golden import RENET_SHAPES, det_global_emb, det_params
    from oracle.stub_global import StubGlobalModel
    from renet_b200 import synthetic
    from renet_b200.model import RENet
    tiny, ev = load_npz('renet_tiny.npz'), load_npz('renet_eval_tiny.npz')
    quads = tiny['quads'].astype(np.int64)
    num_e, R, h, nb = int(tiny['num_e']), int(tiny['R']), int(tiny['h']), int(tiny['nb'])
    seed, num_k = int(ev['seed']), int(ev['num_k'])
    times = np.unique(quads[:, 3])
    params = det_params(RENET_SHAPES(num_e, h, R, nb), seed)
    m = RENet(num_e, h, R, dropout=0, model=0, seq_len=10, num_k=num_k, num_bases=nb)
    m.load_state_dict(params, strict=True)
    m = m.to(device).eval()
    m.global_emb = det_global_emb(times, h, seed + 1)
    m.graph_dict = synthetic.build_graph_dict(quads, R)
    S, ST, O, OT = restate.build_history(quads, num_e)
    pick = lambda L, idx: [L[i] for i in idx]                                           # noqa: E731
    tr, va, te = ev['tr'], ev['va'], ev['te']
    m.init_history(quads[tr], (pick(S, tr), pick(ST, tr)), (pick(O, tr), pick(OT, tr)),
                   quads[va], (pick(S, va), pick(ST, va)), (pick(O, va), pick(OT, va)),
                   quads[te], (pick(S, te), pick(ST, te)), (pick(O, te), pick(OT, te)))
    gm = StubGlobalModel(num_e, h, seed + 2)
    return dict(model=m, ev=ev, quads=quads, hist=(S, ST, O, OT), gm=gm, params=params, dims=(num_e, R, h, nb),
                t_test=int(quads[te[0], 3]))


def eval_flow(ctx, device):
    """Replays oracle/gen_golden.gen_renet_eval_tiny on the product model; returns the same keys."""
    m, ev, quads, gm = ctx['model'], ctx['ev'], ctx['quads'], ctx['gm']
    S, ST, O, OT = ctx['hist']
    num_e = ctx['dims'][0]
    t_test = ctx['t_test']
    res = {'hist_len_s': np.array([len(x) for x in m.s_hist_test]), 'hist_len_o': np.array([len(x) for x in m.o_hist_test]),
           'hist_last_t_s': np.array([x[-1] if len(x) else -1 for x in m.s_hist_test_t])}
    m.latest_time = torch.tensor(t_test)
    allq = torch.from_numpy(quads).to(device)
    torch.manual_seed(1234)
    out = {k: [] for k in ('raw', 'filt', 'loss', 'sub_pred', 'ob_pred')}
    with torch.no_grad():
        for i in ev['te']:
            trip = torch.from_numpy(quads[i]).to(device)
            sh, oh = (S[i], ST[i]), (O[i], OT[i])
            rolled = int(trip[3]) != int(m.latest_time)
            fr, loss = m.evaluate_filter(trip, sh, oh, gm, allq)
            if rolled:
                res['rolled_at'] = np.int64(i)
                res['after_len_s'] = np.array([len(x) for x in m.s_hist_test])
                res['after_len_o'] = np.array([len(x) for x in m.o_hist_test])
                rows = [np.concatenate([[e], r]) for e in range(num_e)
                        if len(m.s_hist_test_t[e]) and m.s_hist_test_t[e][-1] == t_test and len(m.s_hist_test[e])
                        for r in np.asarray(m.s_hist_test[e][-1]).reshape(-1, 2)]
                res['after_new_s_rows'] = np.unique(np.asarray(rows, dtype=np.int64).reshape(-1, 3), axis=0)
                g = m.graph_dict[t_test]
                res['pred_graph_nodes'] = np.sort(g.ndata['id'].view(-1).numpy())
                res['pred_graph_num_edges'] = np.int64(g.number_of_edges())
            rr, _ = m.evaluate(trip, sh, oh, gm)
            _, sp, op = m.predict(trip, sh, oh, gm)
            out['raw'].append(rr); out['filt'].append(fr); out['loss'].append(loss.item())
            out['sub_pred'].append(sp.cpu().numpy().copy()); out['ob_pred'].append(op.cpu().numpy().copy())
    res.update({k: np.asarray(v) for k, v in out.items()})
    res['gm_calls'] = np.asarray(gm.calls, dtype=np.int64)
    return res


def check_eval_against_golden(res, ev, tol=1e-4):
    for k in ('hist_len_s', 'hist_len_o', 'hist_last_t_s', 'rolled_at', 'after_len_s', 'after_len_o', 'after_new_s_rows',
              'pred_graph_nodes', 'pred_graph_num_edges', 'gm_calls'):
        assert np.array_equal(res[k], ev[k]), k
    for k in ('sub_pred', 'ob_pred', 'loss'):
        assert rel_err(res[k], ev[k]) < tol, (k, rel_err(res[k], ev[k]))
    # ranks are integers / half-integers: exact
    assert np.array_equal(res['raw'], ev['raw'])
    assert np.array_equal(res['filt'], ev['filt'])


def global_setup():
    """Inputs of tests/golden/global_tiny.npz (written by oracle/gen_golden.gen_global_tiny from the unmodified reference):
    the tiny stream's quadruples, deterministic RENet_global parameters, the batch of timestamps and soft targets."""
    from oracle.gen_golden import det_params
    g = load_npz('global_tiny.npz')
    tiny = load_npz('renet_tiny.npz')
    quads = tiny['quads'].astype(np.int64)
    shapes = {str(k): tuple(int(x) for x in v[:2] if x > 0) for k, v in zip(g['shapes_keys'], g['shapes_vals'])}
    params = det_params(shapes, int(g['seed']))
    sel = [int(i) for i in g['t_batch']]
    return dict(g=g, quads=quads, num_e=int(tiny['num_e']), R=int(tiny['R']), h=int(g['h']), params=params, sel=sel,
                times=g['times'].astype(np.int64), t_batch=g['times'].astype(np.int64)[sel],
                tps=g['true_prob_s'], tpo=g['true_prob_o'])
