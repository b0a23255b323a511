"""-m "not gpu": pins the CPU restatement of the global model (oracle/restate.global_*) on outputs of the UNMODIFIED
reference (global_model.py + Aggregator.RGCNAggregator_global; tests/golden/global_tiny.npz), and the host-side logic
of renet_b200.global_model (window selection, whole-graph batching) on that restatement."""
import numpy as np
import pytest
import torch

from helpers import global_setup, rel_err
from oracle import restate

TOL = 5e-6


@pytest.mark.parametrize('pool', [1, 0])
@pytest.mark.parametrize('subj', [True, False])
def test_global_forward_loss_and_gradients(pool, subj):
    c = global_setup()
    g = c['g']
    gd = restate.build_graph_dict(c['quads'], c['R'])
    P = {k: v.clone().requires_grad_(True) for k, v in c['params'].items()}
    loss = restate.global_forward(P, c['t_batch'], c['tps'][c['sel']], c['tpo'][c['sel']], gd, subj, maxpool=pool)
    tag = 'pool%d/%s' % (pool, 'subj' if subj else 'obj')
    assert abs(loss.item() - float(g[tag + '/loss'])) < TOL * abs(float(g[tag + '/loss']))
    loss.backward()
    n = 0
    for k in P:
        key = '%s/grad/%s' % (tag, k)
        if key in g.files:
            assert rel_err(P[k].grad.numpy(), g[key]) < 5e-5, k
            n += 1
    assert n == 11


@pytest.mark.parametrize('pool', [1, 0])
def test_global_predict(pool):
    c = global_setup()
    g = c['g']
    gd = restate.build_graph_dict(c['quads'], c['R'])
    with torch.no_grad():
        s_q, sub = restate.global_predict(c['params'], int(c['times'][7]), gd, True, maxpool=pool)
    assert rel_err(s_q.numpy(), g['pool%d/pred_sq' % pool]) < TOL
    assert rel_err(sub.numpy(), g['pool%d/pred_sub' % pool]) < TOL


def test_global_emb_table():
    """get_global_emb (global_model.py:57-73): global_emb[t_k] = predict(t_{k+1}), the last one one time unit later."""
    c = global_setup()
    g = c['g']
    gd = restate.build_graph_dict(c['quads'], c['R'])
    times = [int(t) for t in c['times']]
    unit = times[1] - times[0]
    np.testing.assert_array_equal(g['pool1/global_emb_keys'], times)
    with torch.no_grad():
        for k, t in enumerate(times):
            nxt = times[k + 1] if k + 1 < len(times) else t + unit
            s_q, _ = restate.global_predict(c['params'], nxt, gd, True, maxpool=1)
            assert rel_err(s_q.numpy(), g['pool1/global_emb'][k]) < TOL, t


def test_host_windows_and_batching_match_restatement():
    """renet_b200.global_model host logic (no kernels): windows per timestamp, CSR of the batched whole graphs."""
    from renet_b200 import synthetic
    from renet_b200.global_model import RGCNAggregator_global, whole_graph_arrays
    c = global_setup()
    gd_ref = restate.build_graph_dict(c['quads'], c['R'])
    gd = synthetic.build_graph_dict(c['quads'], c['R'])
    agg = RGCNAggregator_global(c['h'], 0.0, c['num_e'], c['R'], 100, 0, seq_len=10, maxpool=1)
    t_sorted = np.sort(c['t_batch'])[::-1].copy()
    windows, lens = agg._windows(torch.from_numpy(t_sorted), gd)
    want = restate.global_windows(t_sorted, list(gd_ref.keys()), 10)
    assert [list(map(int, w)) for w in windows] == [list(map(int, w)) for w in want]
    assert lens == [len(w) for w in want]
    uniq = sorted({int(t) for w in want for t in w})
    node_ent, norm, rp, col_src, col_ts, col_to, sizes, seg = whole_graph_arrays([gd[t] for t in uniq])
    gs = [gd_ref[t] for t in uniq]
    off = np.concatenate(([0], np.cumsum([x.number_of_nodes() for x in gs])))
    np.testing.assert_array_equal(seg, off)
    src = np.concatenate([x.src + o for x, o in zip(gs, off[:-1])])
    dst = np.concatenate([x.dst + o for x, o in zip(gs, off[:-1])])
    ts = np.concatenate([x.type_s for x in gs])
    to = np.concatenate([x.type_o for x in gs])
    got = sorted(zip(np.repeat(np.arange(len(rp) - 1), np.diff(rp)).tolist(), col_src.tolist(), col_ts.tolist(), col_to.tolist()))
    assert got == sorted(zip(dst.tolist(), src.tolist(), ts.tolist(), to.tolist()))
    np.testing.assert_array_equal(node_ent, np.concatenate([x.id for x in gs]))
    np.testing.assert_array_equal(norm, np.concatenate([x.norm for x in gs]))
