"""not gpu: host logic of the test-time path (renet_b200/inference.py) against the reference's golden run.

The CUDA encode is replaced by the CPU oracle here (tests may do that; the product may not), so what this pins is
init_history, the roll-over control flow (sampling, candidate selection, caches, predicted graph, history roll), the
decoders and the rank rules.  tests/test_gpu_inference.py runs the same flow on the kernels."""
import numpy as np
import torch

from helpers import check_eval_against_golden, eval_flow, eval_setup
from oracle import restate


def _plain(g):
    if isinstance(g, restate.PlainGraph):
        return g
    src, dst, ts, to = g._coo
    return restate.PlainGraph(g.node_id, src.astype(np.int64), dst.astype(np.int64), ts.astype(np.int64), to.astype(np.int64))


def _oracle_encode(ctx):
    params, (num_e, R, h, nb) = ctx['params'], ctx['dims']

    def encode(hist, s, r, ent_embeds, rel_embeds, graph_dict, global_emb, reverse, encoder, encoder_r):
        s, r = np.asarray(s.cpu()).reshape(-1), np.asarray(r.cpu()).reshape(-1)
        z = np.zeros_like(s)
        tr = np.stack((z, r, s), 1) if reverse else np.stack((s, r, z), 1)
        gd = {t: _plain(g) for t, g in graph_dict.items()}
        out = restate.renet_forward(params, tr, hist[0], hist[1], gd, global_emb, not reverse, R, nb)
        assert np.array_equal(out['batch'].s_idx, np.arange(len(s)))     # equal lengths: stable sort = identity
        return out['s_h'], out['s_q'], None
    return encode


def test_rank_rule_and_cache_update():
    from renet_b200.inference import history_triples, rank_with_ties
    sc = torch.tensor([0.5, 0.9, 0.5, 0.1, 0.5])
    assert rank_with_ties(sc, 1) == 1 and rank_with_ties(sc, 3) == 5
    assert rank_with_ties(sc, 0) == 1 + (3 - 1) / 2 + 1                  # model.py:373-379 tie rule
    from renet_b200.model import RENet
    m = RENet(10, 4, 3, num_bases=2)
    c = m.update_cache([], torch.tensor(2), torch.tensor([[3], [13]]))   # % in_dim (model.py:422)
    assert c.tolist() == [[2, 3], [2, 3]]
    c = m.update_cache(torch.tensor([[2, 3], [1, 4]]), torch.tensor(2), torch.tensor([[3], [5]]))
    assert c.tolist() == [[2, 3], [1, 4], [2, 5]]
    c = m.update_cache(torch.tensor([[1, 4]]), torch.tensor(2), torch.tensor([[4]]))
    assert c.tolist() == [[1, 4], [2, 4]]
    s_cache = [[] for _ in range(10)]; o_cache = [[] for _ in range(10)]
    s_cache[1] = torch.tensor([[2, 5]]); o_cache[5] = torch.tensor([[2, 1], [0, 7]])
    assert history_triples(s_cache, o_cache).tolist() == [[1, 2, 5], [7, 0, 5]]   # utils.py:95-113


def test_eval_flow_matches_reference_golden_with_oracle_encode():
    ctx = eval_setup('cpu')
    ctx['model'].aggregator.encode = _oracle_encode(ctx)
    res = eval_flow(ctx, 'cpu')
    check_eval_against_golden(res, ctx['ev'])


def test_rebinding_switch_scores_the_triple_itself():
    """RENet.reference_rebinding = False: the first triple after a timestamp change is scored with its own (s, o), i.e.
    exactly like a repeated call (the reference's numbers for that triple come from model.py:279,290 re-binding s / o)."""
    ctx = eval_setup('cpu')
    m, quads, gm = ctx['model'], ctx['quads'], ctx['gm']
    m.aggregator.encode = _oracle_encode(ctx)
    m.reference_rebinding = False
    S, ST, O, OT = ctx['hist']
    i = int(ctx['ev']['rolled_at'])
    m.latest_time = torch.tensor(ctx['t_test'])
    torch.manual_seed(1234)
    trip = torch.from_numpy(quads[i])
    with torch.no_grad():
        l1, sp1, op1 = m.predict(trip, (S[i], ST[i]), (O[i], OT[i]), gm)      # rolls over, then scores (s, o) itself
        l2, sp2, op2 = m.predict(trip, (S[i], ST[i]), (O[i], OT[i]), gm)      # no roll-over
    assert int(m.latest_time) == int(quads[i, 3])
    assert torch.equal(sp1, sp2) and torch.equal(op1, op2) and float(l1) == float(l2)
    # and it is NOT what the reference returns for that call
    k = int(np.flatnonzero(ctx['ev']['te'] == i)[0])
    assert abs(float(l1) - float(ctx['ev']['loss'][k])) > 1e-3


def test_evaluate_stream_metrics_match_reference_ranks():
    """RENet.evaluate_stream (the reference's test.py loop) reproduces MRR / MR / Hits from the reference's own filtered
    ranks of the golden run (test.py:140-150 formulas)."""
    ctx = eval_setup('cpu')
    m, ev, quads, gm = ctx['model'], ctx['ev'], ctx['quads'], ctx['gm']
    m.aggregator.encode = _oracle_encode(ctx)
    S, ST, O, OT = ctx['hist']
    te = ev['te']
    m.latest_time = torch.tensor(ctx['t_test'])
    torch.manual_seed(1234)
    out = m.evaluate_stream(quads[te], ([S[i] for i in te], [ST[i] for i in te]), ([O[i] for i in te], [OT[i] for i in te]),
                            gm, total_data=quads)
    ref = ev['filt'].reshape(-1)
    np.testing.assert_array_equal(out['ranks'], ref)
    assert abs(out['mrr'] - np.mean(1.0 / ref)) < 1e-12 and abs(out['mr'] - np.mean(ref)) < 1e-12
    for k in (1, 3, 10):
        assert out['hits@%d' % k] == float(np.mean(ref <= k))
    assert abs(out['loss'] - float(ev['loss'].sum())) < 1e-3 * float(ev['loss'].sum())
