"""-m gpu: the device batcher (renet_host_plan_batch + renet_induce_edges) builds, bit for bit, the batch of the
all-host C++ batcher (itself pinned on the numpy path and the oracle in tests/test_host_batching.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _stores(num_timestamps, seed):
    from renet_b200 import hoststore, synthetic
    quads, num_e, num_r = synthetic.make_quads('icews18', seed=seed, num_timestamps=num_timestamps)
    S, ST, O, OT = synthetic.build_history(quads)
    gs = hoststore.GraphStore(synthetic.build_graph_dict(quads, num_r))
    return quads, gs, hoststore.HistoryStore(S, ST, quads[:, 0], gs), hoststore.HistoryStore(O, OT, quads[:, 2], gs)


@pytest.mark.parametrize('num_timestamps,batch', [(16, 300), (60, 1024)])
def test_device_batcher_equals_host_batcher(num_timestamps, batch):
    from renet_b200 import hoststore
    quads, gs, hs_s, hs_o = _stores(num_timestamps, 11)
    rng = np.random.RandomState(3)
    for hs in (hs_s, hs_o):
        for trial in range(2):
            sel = rng.permutation(len(quads))[:batch] if trial == 0 else np.arange(len(quads) - batch, len(quads))
            view = hs.select(sel)
            a = hoststore.assemble_view(view, DEV, device_edges=False)
            b = hoststore.assemble_view(view, DEV, device_edges=True)
            ga, gb = a.graph, b.graph
            assert gb._E is None and gb.E_cap >= ga.E          # the count is still in flight ...
            assert gb.E == ga.E and gb.N == ga.N                # ... and resolves to the host batcher's
            for k in ('node_ent', 'row_ptr', 'col_src', 'col_type_s', 'col_type_o', 'norm', 'seq_len_dev'):
                assert torch.equal(getattr(ga, k), getattr(gb, k)), k
            for k in ('readout', 'row_glob', 'row_seq', 'seq_start', 'packed_row', 's_idx_dev', 'comp_graph_dev'):
                assert torch.equal(getattr(a, k), getattr(b, k)), k
            np.testing.assert_array_equal(a.s_idx, b.s_idx)
            np.testing.assert_array_equal(a.batch_sizes, b.batch_sizes)
            np.testing.assert_array_equal(a.times, b.times)
            np.testing.assert_array_equal(a.seq_len, b.seq_len)


def test_device_batcher_degenerate_cases():
    """all-empty batch; a batch whose induced graph keeps no edge (single sample, history of isolated pairs is not
    constructible from real data, so emulate with direct renet_induce_edges calls)."""
    from renet_b200 import _lib, hoststore
    quads, gs, hs_s, _ = _stores(8, 5)
    hb = hoststore.assemble_view(hs_s.select(np.asarray([0, 1])), DEV, device_edges=True)
    assert hb.graph is None and hb.S == 0
    L = _lib.lib()
    P = _lib.ptr
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=DEV)      # noqa: E731
    # one timestamp graph with 3 nodes, edges (dst-sorted) 1->0, 2->0, 0->1, 0->2 ; select nodes {1,2}: nothing survives
    edge_off = torch.tensor([0, 4], dtype=torch.int64, device=DEV)
    src, dst = i32([1, 2, 0, 0]), i32([0, 0, 1, 2])
    ts, to = i32([5, 6, 7, 8]), i32([15, 16, 17, 18])
    for newid, want_rp, want_src in (([-1, 0, 1], [0, 0, 0], []), ([0, -1, 1], [0, 1, 2], [1, 0]), ([0, 1, 2], [0, 2, 3, 4], [1, 2, 0, 0])):
        N = sum(1 for x in newid if x >= 0)
        nid = i32(newid)
        row_ptr = torch.full((N + 1,), -7, dtype=torch.int32, device=DEV)
        cs, cts, cto = (torch.full((4,), -7, dtype=torch.int32, device=DEV) for _ in range(3))
        norm = torch.zeros(N, device=DEV)
        ec = i32([-1])
        ws = torch.empty(int(L.renet_induce_workspace_bytes(4)) // 4, dtype=torch.int32, device=DEV)
        rc = L.renet_induce_edges(P(edge_off), P(src), P(dst), P(ts), P(to), P(i32([0])), P(i32([0, 3])), P(i32([0, 4])), P(nid),
                                  1, N, 4, P(row_ptr), P(cs), P(cts), P(cto), P(norm), P(ec), P(ws), ws.numel() * 4, _lib.stream())
        _lib.check(rc, 'renet_induce_edges')
        E = int(ec.item())
        assert E == len(want_src) and row_ptr.tolist() == want_rp and cs[:E].tolist() == want_src
        deg = np.diff(np.asarray(want_rp))
        np.testing.assert_array_equal(norm.cpu().numpy(), (1.0 / np.maximum(deg, 1)).astype(np.float32))


def test_encode_identical_on_both_batchers():
    """RENet.encode gives the same GRU states from a device-assembled and a host-assembled batch."""
    from renet_b200 import hoststore
    from renet_b200.model import RENet
    quads, gs, hs_s, hs_o = _stores(16, 7)
    torch.manual_seed(0)
    m = RENet(23033, 200, 256, num_bases=100).to(DEV).eval()
    m.global_emb = {int(t): torch.randn(200) * 0.1 for t in gs.times}
    sel = np.arange(len(quads) - 256, len(quads))
    batch = torch.from_numpy(quads[sel]).to(DEV)
    outs = []
    for dev_edges in (False, True):
        vs = hoststore.assemble_view(hs_s.select(sel), DEV, device_edges=dev_edges)
        vo = hoststore.assemble_view(hs_o.select(sel), DEV, device_edges=dev_edges)
        with torch.no_grad():
            outs.append([m.encode(batch, vs, vo, gs, subject=subj)[3:5] for subj in (True, False)])
    for (a, b) in zip(outs[0], outs[1]):
        assert torch.allclose(a[0], b[0], atol=1e-6) and torch.allclose(a[1], b[1], atol=1e-6)
