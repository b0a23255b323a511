"""-m gpu: the global model (renet_b200.global_model: RGCN over whole graphs, segment pooling, dense GRU on the tensor-core
engine) against outputs of the UNMODIFIED reference (tests/golden/global_tiny.npz) and the CPU oracle."""
import numpy as np
import pytest
import torch

from helpers import global_setup, rel_err
from oracle import restate

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = 'cuda:0'


def _model(c, pool):
    from renet_b200 import synthetic
    from renet_b200.global_model import RENet_global
    m = RENet_global(c['num_e'], c['h'], c['R'], dropout=0, model=3, seq_len=10, num_k=10, maxpool=pool)
    m.load_state_dict(c['params'], strict=True)          # the reference's state_dict keys load as they are
    return m.to(DEV), synthetic.build_graph_dict(c['quads'], c['R'])


@pytest.mark.parametrize('pool', [1, 0])
@pytest.mark.parametrize('subj', [True, False])
def test_global_forward_backward_matches_reference_golden(pool, subj):
    c = global_setup()
    g = c['g']
    m, gd = _model(c, pool)
    m.train()
    tps = torch.from_numpy(c['tps'][c['sel']]).to(DEV)
    tpo = torch.from_numpy(c['tpo'][c['sel']]).to(DEV)
    loss = m(torch.from_numpy(c['t_batch']), tps, tpo, gd, subject=subj)
    tag = 'pool%d/%s' % (pool, 'subj' if subj else 'obj')
    ref = float(g[tag + '/loss'])
    assert abs(loss.item() - ref) < TOL * abs(ref), (loss.item(), ref)
    loss.backward()
    n = 0
    for k, p in m.named_parameters():
        key = '%s/grad/%s' % (tag, k)
        if key in g.files:
            assert p.grad is not None, k
            assert rel_err(p.grad.cpu().numpy(), g[key]) < 5e-4, k
            n += 1
    assert n == 11


@pytest.mark.parametrize('pool', [1, 0])
def test_global_predict_and_embedding_table(pool):
    c = global_setup()
    g = c['g']
    m, gd = _model(c, pool)
    m.eval()
    with torch.no_grad():
        s_q, sub, prob = m.predict(int(c['times'][7]), gd)
        assert s_q.shape == (1, 1, c['h']) and sub.shape == (1, 1, c['num_e'])
        assert rel_err(s_q.view(-1).cpu().numpy(), g['pool%d/pred_sq' % pool]) < TOL
        assert rel_err(sub.view(-1).cpu().numpy(), g['pool%d/pred_sub' % pool]) < TOL
        assert abs(prob.sum().item() - 1.0) < 1e-5
        if pool == 1:
            ge = m.get_global_emb([int(t) for t in c['times']], gd)
            np.testing.assert_array_equal(sorted(ge), g['pool1/global_emb_keys'])
            got = np.stack([ge[k].view(-1).cpu().numpy() for k in sorted(ge)])
            assert rel_err(got, g['pool1/global_emb']) < TOL
            packed = m.aggregator(torch.from_numpy(c['times'][[12, 5, 3]]), m.ent_embeds, gd, reverse=False)
            np.testing.assert_array_equal(packed.batch_sizes.numpy(), g['pool1/agg_bs'])
            assert rel_err(packed.data.cpu().numpy(), g['pool1/agg_packed']) < TOL


def test_segment_pool_kernels_vs_torch():
    """renet_segment_pool_fwd/_bwd: ragged segments incl. single-row and large ones, ties broken towards the first row
    (torch.max semantics are unspecified on ties; gradients are compared on tie-free data)."""
    from renet_b200.global_model import _SegmentPoolFn
    gen = torch.Generator().manual_seed(3)
    sizes = [1, 7, 300, 2, 1025, 64]
    off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)
    H = torch.randn(int(off[-1]), 200, generator=gen)
    G = torch.randn(len(sizes), 200, generator=gen)
    seg = torch.from_numpy(off).to(DEV)
    for mode in (1, 0):
        Hd = H.to(DEV).requires_grad_(True)
        out = _SegmentPoolFn.apply(Hd, seg, mode)
        (out * G.to(DEV)).sum().backward()
        Hc = H.clone().requires_grad_(True)
        rows = [Hc[a:b].max(0).values if mode == 1 else Hc[a:b].mean(0) for a, b in zip(off[:-1], off[1:])]
        ref = torch.stack(rows)
        (ref * G).sum().backward()
        assert rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < 1e-6
        assert rel_err(Hd.grad.cpu().numpy(), Hc.grad.numpy()) < 1e-6


def test_dense_gru_vs_oracle():
    """renet_gru_dense_fwd/_bwd (single GRU(k,h)) against the CPU restatement incl. gradients, ragged lengths."""
    from renet_b200.global_model import gru_final_hidden
    gen = torch.Generator().manual_seed(5)
    h, k = 200, 200
    lens = [10, 10, 9, 7, 7, 4, 2, 1, 1]
    S = sum(lens)
    X = torch.randn(S, k, generator=gen) * 0.5
    gru = torch.nn.GRU(k, h, batch_first=True)
    G = torch.randn(len(lens), h, generator=gen)
    P = [p.detach().clone().requires_grad_(True) for p in (gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0)]
    Xc = X.clone().requires_grad_(True)
    ref = restate.gru_final_hidden_batched(Xc, lens, *P)
    (ref * G).sum().backward()
    gd = gru.to(DEV)
    Xd = X.to(DEV).requires_grad_(True)
    out = gru_final_hidden(gd, Xd, lens)
    (out * G.to(DEV)).sum().backward()
    assert rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < TOL
    assert rel_err(Xd.grad.cpu().numpy(), Xc.grad.numpy()) < 5e-4
    for p, q in zip((gd.weight_ih_l0, gd.weight_hh_l0, gd.bias_ih_l0, gd.bias_hh_l0), P):
        assert rel_err(p.grad.cpu().numpy(), q.grad.numpy()) < 5e-4
