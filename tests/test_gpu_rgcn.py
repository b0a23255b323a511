"""-m gpu: the CUDA RGCN block layer (through the C-ABI) against the reference's golden vectors and the
CPU oracle.  Tolerance: 1e-4 relative (max-abs-diff / max-abs-ref), the north-star bar; integer-valued
known-answer cases must be exact."""
import numpy as np
import pytest
import torch

from helpers import layer_case, load_npz, rel_err, t
from oracle import restate

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope='module')
def G():
    import gpu_helpers
    assert torch.cuda.is_available()
    return gpu_helpers


def _case_tensors(G, c):
    et = c['type_o'] if int(c['reverse']) else c['type_s']
    H, W = G.d(c['H']), G.d(c['W'])
    Wl = G.d(c['Wloop']) if bool(c['self_loop']) else None
    return H, W, Wl, et


def test_golden_layer_cases_forward(G):
    blob = load_npz('layer_cases.npz')
    for name in blob['names']:
        c = layer_case(blob, str(name))
        N, E = int(c['N']), len(c['src'])
        H, W, Wl, et = _case_tensors(G, c)
        rp, cs, ct = G.csr_from_coo(c['src'], c['dst'], et, N)
        out = G.layer_fwd(H, None, W, Wl, rp, cs, ct, G.d(c['ref_norm']), N, E, int(c['d_in']), int(c['d_out']),
                          int(c['nb']), bool(c['relu']))
        err = rel_err(out.cpu().numpy(), c['ref_out'])
        if str(name) in ('hand_kat', 'dup_edge', 'reverse'):
            assert err == 0.0, (name, err)          # small integers: exact in fp32
        assert err < TOL, (name, err)


def test_golden_layer_cases_backward(G):
    blob = load_npz('layer_cases.npz')
    for name in blob['names']:
        c = layer_case(blob, str(name))
        N, E = int(c['N']), len(c['src'])
        H, W, Wl, et = _case_tensors(G, c)
        rp, cs, ct = G.csr_from_coo(c['src'], c['dst'], et, N)
        norm = G.d(c['ref_norm'])
        out = G.layer_fwd(H, None, W, Wl, rp, cs, ct, norm, N, E, int(c['d_in']), int(c['d_out']), int(c['nb']),
                          bool(c['relu']))
        dH, dW, dWl = G.layer_bwd(H, None, W, Wl, c['src'], c['dst'], et, norm, out, G.d(c['G']), N, E,
                                  int(c['d_in']), int(c['d_out']), int(c['nb']), bool(c['relu']))
        assert rel_err(dH.cpu().numpy(), c['ref_dH']) < TOL, name
        assert rel_err(dW.cpu().numpy(), c['ref_dW']) < TOL, name
        if Wl is not None:
            assert rel_err(dWl.cpu().numpy(), c['ref_dWloop']) < TOL, name


def test_zero_edge_graph(G):
    blob = load_npz('layer_cases.npz')
    H, W, Wl = G.d(blob['zero_edge/H']), G.d(blob['zero_edge/W']), G.d(blob['zero_edge/Wloop'])
    rp = torch.zeros(5, dtype=torch.int32, device=G.DEV)
    norm = torch.ones(4, device=G.DEV)
    out = G.layer_fwd(H, None, W, Wl, rp, None, None, norm, 4, 0, 4, 4, 2, True)
    assert rel_err(out.cpu().numpy(), blob['zero_edge/ref_out']) < 1e-6


def test_empty_graph_and_bad_args(G):
    from renet_b200 import _lib
    L = _lib.lib()
    assert L.renet_rgcn_block_fwd(None, None, None, None, None, None, None, None, None, 0, 0, 200, 200, 100, 8, 1, None) == 0
    W = torch.zeros(8, 400, device=G.DEV)
    rc = L.renet_rgcn_block_fwd(_lib.ptr(W), None, _lib.ptr(W), None, None, None, None, None, None, 4, 3, 200, 200, 100, 8, 1, None)
    assert rc == -1


@pytest.fixture(scope='module')
def icews_batch():
    from renet_b200 import synthetic, utils
    tkg = synthetic.SyntheticTKG('icews18', seed=999, num_timestamps=40)
    q, sh, oh = tkg.batch(0, batch_size=1024, tail_only=True)
    hb = utils.assemble_history_batch_host(sh[0], sh[1], q[:, 0], tkg.graph_dict)
    return tkg, hb


def _oracle_layer(hb, H, W, Wl, et, relu):
    g = hb.graph
    dst = np.repeat(np.arange(len(g['node_ent'])), np.diff(g['row_ptr']))
    return restate.rgcn_block_layer(H, W, Wl, t(g['col_src']), t(dst), t(et.astype(np.int64)), t(g['norm']), relu, 100)


def test_icews18_shaped_batch_two_layers_vs_oracle(G, icews_batch):
    """Full-size ICEWS18-shaped history graph (about 11 k nodes / 33 k edges per 40-timestamp stream
    batch; the 240-timestamp bench shape is covered by the property tests below): both layers, the
    fused embedding lookup (h_index) included, against the CPU oracle on the same seeded inputs."""
    tkg, hb = icews_batch
    g = hb.graph
    N, E = len(g['node_ent']), len(g['col_src'])
    torch.manual_seed(0)
    ent = torch.randn(tkg.num_e, 200) * 0.1
    W1, W2 = torch.randn(512, 400) * 0.1, torch.randn(512, 400) * 0.1
    L1, L2 = torch.randn(200, 200) * 0.07, torch.randn(200, 200) * 0.07
    for reverse, et in ((False, g['col_type_s']), (True, g['col_type_o'])):
        H0 = ent[t(g['node_ent'])]
        o1 = _oracle_layer(hb, H0, W1, L1, et, True)
        o2 = _oracle_layer(hb, o1, W2, L2, et, False)
        rp, cs, ct = G.d(g['row_ptr'], torch.int32), G.d(g['col_src'], torch.int32), G.d(et, torch.int32)
        norm, idx = G.d(g['norm']), G.d(g['node_ent'], torch.int32)
        h1 = G.layer_fwd(ent.to(G.DEV), idx, W1.to(G.DEV), L1.to(G.DEV), rp, cs, ct, norm, N, E, 200, 200, 100, True)
        h2 = G.layer_fwd(h1, None, W2.to(G.DEV), L2.to(G.DEV), rp, cs, ct, norm, N, E, 200, 200, 100, False)
        assert rel_err(h1.cpu().numpy(), o1.numpy()) < TOL
        assert rel_err(h2.cpu().numpy(), o2.numpy()) < TOL


def test_module_autograd_vs_oracle(G, icews_batch):
    """RGCNBlockLayer modules (autograd.Function over the CUDA kernels) vs torch autograd on the oracle."""
    import torch.nn.functional as F
    from renet_b200 import utils
    from renet_b200.rgcn import RGCNBlockLayer
    tkg, hb_host = icews_batch
    import copy
    hb = utils.upload_history_batch(copy.copy(hb_host), torch.device(G.DEV))
    g = hb.graph
    gh = hb_host.graph if isinstance(hb_host.graph, dict) else None
    torch.manual_seed(1)
    ent = (torch.randn(tkg.num_e, 200) * 0.1)
    l1 = RGCNBlockLayer(200, 200, 512, 100, activation=F.relu, self_loop=True)
    l2 = RGCNBlockLayer(200, 200, 512, 100, activation=None, self_loop=True)
    Gout = torch.randn(g.N, 200)
    # oracle side
    P = [p.detach().clone().requires_grad_(True) for p in (ent, l1.weight, l1.loop_weight, l2.weight, l2.loop_weight)]
    node_ent = g.node_ent.cpu().long()
    dst = torch.repeat_interleave(torch.arange(g.N), (g.row_ptr[1:] - g.row_ptr[:-1]).cpu().long())
    src, et, norm = g.col_src.cpu().long(), g.col_type_s.cpu().long(), g.norm.cpu()
    o1 = restate.rgcn_block_layer(P[0][node_ent], P[1], P[2], src, dst, et, norm, True, 100)
    o2 = restate.rgcn_block_layer(o1, P[3], P[4], src, dst, et, norm, False, 100)
    (o2 * Gout).sum().backward()
    # CUDA side
    l1, l2 = l1.to(G.DEV), l2.to(G.DEV)
    ent_d = ent.to(G.DEV).requires_grad_(True)
    h1 = l1.apply_layer(g, ent_d, g.node_ent, False)
    h2 = l2.apply_layer(g, h1, None, False)
    assert rel_err(h2.detach().cpu().numpy(), o2.detach().numpy()) < TOL
    (h2 * Gout.to(G.DEV)).sum().backward()
    got = [ent_d.grad, l1.weight.grad, l1.loop_weight.grad, l2.weight.grad, l2.loop_weight.grad]
    for a, b, nm in zip(got, P, ('ent', 'W1', 'L1', 'W2', 'L2')):
        assert rel_err(a.cpu().numpy(), b.grad.numpy()) < TOL, nm


def test_full_size_properties(G):
    """Size-independent properties at the bench shape (N ~ 34 k, E ~ 200 k), no oracle needed:
    linearity in H without activation, invariance to the edge order inside the COO list, and
    agreement between host-built and device-built (renet_build_csr) CSR."""
    rng = np.random.RandomState(0)
    N, E, R2 = 34000, 200000, 512
    src, dst = rng.randint(0, N, E), (rng.zipf(1.3, E) % N)
    et = rng.randint(0, R2, E)
    deg = np.bincount(dst, minlength=N).astype(np.float32); deg[deg == 0] = 1
    norm = G.d(1.0 / deg)
    H = torch.randn(N, 200, device=G.DEV)
    W = torch.randn(R2, 400, device=G.DEV) * 0.1
    Wl = torch.randn(200, 200, device=G.DEV) * 0.07
    rp, cs, ct = G.csr_from_coo(src, dst, et, N)
    # device CSR == host CSR (stable)
    order = np.argsort(dst, kind='stable')
    np.testing.assert_array_equal(cs.cpu().numpy(), src[order])
    np.testing.assert_array_equal(ct.cpu().numpy(), et[order])
    np.testing.assert_array_equal(rp.cpu().numpy(), np.concatenate(([0], np.cumsum(np.bincount(dst, minlength=N)))))
    a = G.layer_fwd(H, None, W, Wl, rp, cs, ct, norm, N, E, 200, 200, 100, False)
    b = G.layer_fwd(H * 3.0, None, W, Wl, rp, cs, ct, norm, N, E, 200, 200, 100, False)
    assert rel_err(b.cpu().numpy(), (a * 3.0).cpu().numpy()) < 1e-5
    perm = rng.permutation(E)
    rp2, cs2, ct2 = G.csr_from_coo(src[perm], dst[perm], et[perm], N)
    c = G.layer_fwd(H, None, W, Wl, rp2, cs2, ct2, norm, N, E, 200, 200, 100, False)
    assert rel_err(c.cpu().numpy(), a.cpu().numpy()) < 1e-5
    # the gather hands partial sums between warps in a fixed order (no atomics): bitwise reproducible, so
    # relu(layer) == max(layer, 0) exactly and a second run gives the same bits
    r = G.layer_fwd(H, None, W, Wl, rp, cs, ct, norm, N, E, 200, 200, 100, True)
    assert torch.equal(r, torch.clamp_min(a, 0))
    assert torch.equal(G.layer_fwd(H, None, W, Wl, rp, cs, ct, norm, N, E, 200, 200, 100, False), a)


def test_component_resident_kernel_vs_oracle(G):
    """renet_rgcn_gather_comp on a batched graph with components larger than the shared-memory window
    (144 rows), with and without hot relations / launch order, against the CPU oracle and the tile kernel."""
    from renet_b200 import _lib, utils
    L = _lib.lib()
    rng = np.random.RandomState(5)
    sizes = [300, 17, 144, 145, 1, 64, 250]
    comp_start = np.concatenate(([0], np.cumsum(sizes)))
    N, R2 = int(comp_start[-1]), 96
    src, dst, comp_edges = [], [], []
    for c, n in enumerate(sizes):
        e = n * 7 if n > 1 else 3
        src.append(comp_start[c] + rng.randint(0, n, e)); dst.append(comp_start[c] + rng.randint(0, n, e))
        comp_edges.append(e)
    src, dst = np.concatenate(src), np.concatenate(dst)
    order = np.argsort(dst, kind='stable')
    src, dst = src[order], dst[order]
    et = (rng.zipf(1.4, len(src)) % R2).astype(np.int64)
    indeg = np.bincount(dst, minlength=N)
    row_ptr = np.concatenate(([0], np.cumsum(indeg)))
    norm = (1.0 / np.maximum(indeg, 1)).astype(np.float32)
    comp_edges = np.bincount(np.searchsorted(comp_start[1:], dst, side='right'), minlength=len(sizes))
    ex = utils.component_extras(comp_start, comp_edges, et, et, num_types=R2)
    torch.manual_seed(2)
    ent = torch.randn(500, 200) * 0.3
    node_ent = rng.randint(0, 500, N)
    W, Wl = torch.randn(R2, 400) * 0.1, torch.randn(200, 200) * 0.07
    ref = restate.rgcn_block_layer(ent[t(node_ent)], W, Wl, t(src), t(dst), t(et), t(norm), True, 100)
    d = lambda a, dt=torch.int32: G.d(np.asarray(a), dt)
    rp, cs, ct, nrm, idx = d(row_ptr), d(src), d(et), G.d(norm), d(node_ent)
    Wd, Wld, entd = W.to(G.DEV), Wl.to(G.DEV), ent.to(G.DEV)
    tile = G.layer_fwd(entd, idx, Wd, Wld, rp, cs, ct, nrm, N, len(src), 200, 200, 100, True)
    assert rel_err(tile.cpu().numpy(), ref.numpy()) < TOL
    cptr_d, cord_d, slot_d, hot_d = d(ex['comp_ptr']), d(ex['comp_order']), d(ex['rel_slot_s']), d(ex['hot_s'])
    for use_hot, use_order in ((True, True), (False, False), (True, False)):
        out = torch.empty(N, 200, device=G.DEV)
        _lib.check(L.renet_selfloop_gemm(_lib.ptr(entd), _lib.ptr(idx), _lib.ptr(Wld), _lib.ptr(out), N, 200, 200,
                                         _lib.stream()), 'gemm')
        rc = L.renet_rgcn_gather_comp(_lib.ptr(entd), _lib.ptr(idx), _lib.ptr(Wd), _lib.ptr(rp), _lib.ptr(cs),
                                      _lib.ptr(ct), _lib.ptr(nrm), _lib.ptr(out), _lib.ptr(cptr_d),
                                      _lib.ptr(cord_d) if use_order else None,
                                      _lib.ptr(slot_d) if use_hot else None,
                                      _lib.ptr(hot_d) if use_hot else None, ex['n_hot_s'] if use_hot else 0,
                                      N, len(src), len(sizes), 200, 200, 100, R2, 1, 1, _lib.stream())
        _lib.check(rc, 'renet_rgcn_gather_comp')
        assert rel_err(out.cpu().numpy(), ref.numpy()) < TOL, (use_hot, use_order)
    # every forward variant gives the same result; the ring (6) and hot-relation (7) kernels keep the default kernel's
    # summation order and are bit-identical to it
    hot = np.ascontiguousarray(np.argsort(-np.bincount(et, minlength=R2), kind='stable')[:40].astype(np.int32))
    _lib.check(L.renet_set_hot_relations(hot.ctypes.data_as(_lib.ctypes.c_void_p), len(hot), R2), 'hot')
    try:
        for variant in (1, 2, 6, 7, 0):
            L.renet_set_gather_variant(variant)
            other = G.layer_fwd(entd, idx, Wd, Wld, rp, cs, ct, nrm, N, len(src), 200, 200, 100, True)
            assert rel_err(other.cpu().numpy(), ref.numpy()) < TOL, variant
            if variant in (6, 7, 0):
                assert torch.equal(other, tile), variant
    finally:
        L.renet_set_gather_variant(0)
        L.renet_set_hot_relations(None, 0, 0)
