"""GPU parity of the sm_100a encoder (through the C-ABI) against the golden fixtures and the C oracle."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _data(c, device='cpu'):
    from esc_gnn_b200.data import Data
    d = Data(x=torch.ones(c.n, 1), edge_index=torch.as_tensor(c.ei, dtype=torch.long, device=device))
    d.num_nodes = c.n
    return d


@pytest.mark.parametrize('fname', G.FILES)
def test_create_subgraphs_host_path_matches_reference_fixtures(fname):
    from esc_gnn_b200.transform import create_subgraphs
    for c in G.load(fname):
        o = create_subgraphs(_data(c), c.h, use_rd=c.use_rd, self_loop=c.self_loop)
        assert o.pos_enc.dtype == o.pos_index.dtype == o.pos_batch.dtype == torch.int64
        G.check_against_case(c, o.edge_index.numpy(), o.pos_enc.numpy(), o.pos_index.numpy(), o.pos_batch.numpy())
        if c.use_rd:     # rd block: bit-exact against the fp64 oracle (parity policy E5)
            ref = c_oracle.encode_graph(c.ei, c.n, c.h, True, c.self_loop)
            for got, want in zip((o.edge_index, o.pos_enc, o.pos_index, o.pos_batch), ref):
                assert np.array_equal(got.numpy(), want), c.name


@pytest.mark.parametrize('fname', ('kat1', 'edge_cases', 'cfg2', 'cfg4'))
def test_create_subgraphs_device_path(fname):
    from esc_gnn_b200.transform import create_subgraphs
    for c in G.load(fname)[:30]:
        o = create_subgraphs(_data(c, 'cuda'), c.h, use_rd=c.use_rd, self_loop=c.self_loop)
        assert o.pos_enc.is_cuda
        ref = c_oracle.encode_graph(c.ei, c.n, c.h, c.use_rd, c.self_loop)
        for got, want in zip((o.edge_index, o.pos_enc, o.pos_index, o.pos_batch), ref):
            assert np.array_equal(got.cpu().numpy(), want), c.name


def _oracle_batch(src, dst, eptr, nptr, h, use_rd, self_loop):
    eo, pe, pi, pb = [], [], [], []
    base = 0
    for g in range(len(nptr) - 1):
        a, b = eptr[g], eptr[g + 1]
        r = c_oracle.encode_graph(np.stack([src[a:b], dst[a:b]]), int(nptr[g + 1] - nptr[g]), h, use_rd, self_loop)
        eo.append(r[0]); pe.append(r[1]); pi.append(r[2]); pb.append(r[3] + base)
        base += r[0].shape[1]
    return np.concatenate(eo, 1), np.concatenate(pe), np.concatenate(pi), np.concatenate(pb)


@pytest.mark.parametrize('config,count', [(1, 96), (2, 256), (3, 32), (4, 64)])
def test_batched_encoder_matches_oracle_on_config_shapes(config, count):
    from esc_gnn_b200 import synth
    from esc_gnn_b200.transform import encode_batch, encode_batch_host
    fl = synth.ENCODER_FLAGS[config]
    src, dst, eptr, nptr = synth.make_batch_arrays(config, 5000, count)
    want = _oracle_batch(src, dst, eptr, nptr, fl['h'], fl['use_rd'], fl['self_loop'])
    r = encode_batch(torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda(), torch.as_tensor(eptr),
                     torch.as_tensor(nptr), fl['h'], fl['use_rd'], fl['self_loop'])
    for got, w in zip((r.edge_index, r.pos_enc, r.pos_index, r.pos_batch), want):
        assert np.array_equal(got.cpu().numpy(), w)
    rh = encode_batch_host(src, dst, eptr, nptr, fl['h'], fl['use_rd'], fl['self_loop'])
    for got, w in zip((rh.edge_index, rh.pos_enc, rh.pos_index, rh.pos_batch), want):
        assert np.array_equal(got.numpy(), w)
    # compact records agree with the expanded triple
    rec = r.rec.to(torch.int64) & 0xffffffff
    k = 0
    off, nz = r.rec_off.cpu().numpy(), r.rec_nnz.cpu().numpy()
    recs = rec.cpu().numpy()
    for e in (0, len(off) // 2, len(off) - 1):
        seg = recs[off[e]:off[e] + nz[e]]
        m = want[3] == e
        assert np.array_equal(seg & 2047, want[2][m]) and np.array_equal(seg >> 11, want[1][m])


@pytest.mark.parametrize('h,self_loop', [(1, False), (2, True), (3, False), (4, True)])
def test_sweep_shapes_match_oracle(h, self_loop):
    """Config 5 prefix: graphs of 25..500 nodes (shared-memory path with the 4-bit distance matrix)."""
    from esc_gnn_b200 import synth
    from esc_gnn_b200.transform import encode_batch
    src, dst, eptr, nptr = synth.make_batch_arrays(5, 100 * h, 48)
    want = _oracle_batch(src, dst, eptr, nptr, h, False, self_loop)
    r = encode_batch(torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda(), torch.as_tensor(eptr),
                     torch.as_tensor(nptr), h, False, self_loop)
    for got, w in zip((r.edge_index, r.pos_enc, r.pos_index, r.pos_batch), want):
        assert np.array_equal(got.cpu().numpy(), w)


def test_large_graph_uses_global_slab_and_matches_oracle():
    from esc_gnn_b200 import synth
    from esc_gnn_b200.transform import encode_batch
    rng = np.random.Generator(np.random.PCG64(7))
    parts = []
    for n in (1500, 40, 900):
        und = synth.random_graph(rng, n, int(1.2 * n))
        parts.append((synth.symmetrise(und), n))
    src = np.concatenate([p[0][0] for p in parts]); dst = np.concatenate([p[0][1] for p in parts])
    eptr = np.cumsum([0] + [p[0].shape[1] for p in parts]); nptr = np.cumsum([0] + [p[1] for p in parts])
    want = _oracle_batch(src, dst, eptr, nptr, 2, False, True)
    r = encode_batch(torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda(), torch.as_tensor(eptr),
                     torch.as_tensor(nptr), 2, False, True)
    for got, w in zip((r.edge_index, r.pos_enc, r.pos_index, r.pos_batch), want):
        assert np.array_equal(got.cpu().numpy(), w)


def test_rd_cta_mode_matches_oracle():
    """Graphs too large for the warp-per-edge solver (n > ~40) go through the CTA-per-edge solver."""
    from esc_gnn_b200 import synth
    from esc_gnn_b200.transform import encode_batch
    rng = np.random.Generator(np.random.PCG64(11))
    parts = []
    for n in (110, 20, 75, 130):
        und = synth.random_graph(rng, n, n + 3, max_degree=4)
        parts.append((synth.symmetrise(und), n))
    src = np.concatenate([p[0][0] for p in parts]); dst = np.concatenate([p[0][1] for p in parts])
    eptr = np.cumsum([0] + [p[0].shape[1] for p in parts]); nptr = np.cumsum([0] + [p[1] for p in parts])
    for h, sl in ((3, False), (4, True)):
        want = _oracle_batch(src, dst, eptr, nptr, h, True, sl)
        r = encode_batch(torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda(), torch.as_tensor(eptr),
                         torch.as_tensor(nptr), h, True, sl)
        for got, w in zip((r.edge_index, r.pos_enc, r.pos_index, r.pos_batch), want):
            assert np.array_equal(got.cpu().numpy(), w)


def test_invariants_at_full_batch_size():
    """SURVEY KAT-3 on a BASELINE-sized batch (8192 ZINC-shaped graphs): size-independent properties."""
    from esc_gnn_b200 import synth
    from esc_gnn_b200.transform import encode_batch
    pool = synth.make_batch_arrays(2, 0, 512)
    reps = 16
    src = np.tile(pool[0], reps); dst = np.tile(pool[1], reps)
    eptr = np.concatenate([[0], np.cumsum(np.tile(np.diff(pool[2]), reps))])
    nptr = np.concatenate([[0], np.cumsum(np.tile(np.diff(pool[3]), reps))])
    r = encode_batch(torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda(), torch.as_tensor(eptr),
                     torch.as_tensor(nptr), 3, True, False)
    E = r.num_edges
    assert E == len(src)
    pe, pi, pb = r.pos_enc, r.pos_index, r.pos_batch
    assert bool((pb[1:] >= pb[:-1]).all())                                   # edges in order
    same = pb[1:] == pb[:-1]
    assert bool((pi[1:][same] > pi[:-1][same]).all())                        # ascending index inside an edge
    def block_sum(lo, hi):
        m = (pi >= lo) & (pi < hi)
        return torch.zeros(E, dtype=torch.int64, device=pe.device).index_add_(0, pb[m], pe[m])
    s_deg, s_d0, s_d1, s_rd = block_sum(0, 200), block_sum(200, 300), block_sum(300, 400), block_sum(400, 500)
    assert bool((s_deg == s_d0).all() and (s_d0 == s_d1).all() and (s_d1 == s_rd).all())   # each = #S
    one = torch.zeros(E, dtype=torch.int64, device=pe.device).index_add_(0, pb[pi == 200], pe[pi == 200])
    assert bool((one == 1).all())                                            # exactly one node at d0 = 0 (u != v)
    # periodicity: replica k equals replica 0
    n0 = int((pb < pool[2][-1]).sum())
    assert pe.numel() == n0 * reps
    assert torch.equal(pe[:n0], pe[n0 * (reps - 1):]) and torch.equal(pi[:n0], pi[n0 * (reps - 1):])


def test_error_behaviour_matches_reference():
    from esc_gnn_b200.data import Data
    from esc_gnn_b200.transform import create_subgraphs
    hub = torch.tensor([[0] * 210 + list(range(1, 211)), list(range(1, 211)) + [0] * 210])
    d = Data(x=torch.ones(211, 1), edge_index=hub)
    with pytest.raises(RuntimeError):
        create_subgraphs(d, 1)                                   # degree >= 200: one_hot raises in the reference
    with pytest.raises(RuntimeError):
        create_subgraphs(Data(x=torch.ones(3, 1), edge_index=hub[:, :2]), 5)        # h >= 5
    with pytest.raises(RuntimeError):
        create_subgraphs(Data(x=torch.ones(3, 1), edge_index=torch.zeros((2, 0), dtype=torch.long)), 2)  # cat([])
    with pytest.raises(RuntimeError):
        create_subgraphs(Data(x=torch.ones(3, 1), edge_index=torch.tensor([[0, 1], [1, 2]])), 2, use_rd=True)
    with pytest.raises(RuntimeError):
        create_subgraphs(Data(x=torch.ones(2, 1), edge_index=torch.tensor([[0, 5], [1, 0]])), 2)         # bad id


def test_edge_attr_self_loop_rewrite():
    """E1 on attributes: loop rows dropped, N rows of 1.0 appended (PyG add_self_loops, utils_edge_efficient.py:35-36)."""
    from esc_gnn_b200.data import Data
    from esc_gnn_b200.transform import create_subgraphs
    ei = torch.tensor([[0, 0, 1, 1, 2], [0, 1, 0, 2, 1]])
    ea = torch.arange(10, dtype=torch.float32).view(5, 2)
    o = create_subgraphs(Data(x=torch.ones(3, 1), edge_index=ei, edge_attr=ea, y=torch.tensor([1.0])), 2, self_loop=True)
    assert o.edge_index.tolist() == [[0, 1, 1, 2, 0, 1, 2], [1, 0, 2, 1, 0, 1, 2]]
    assert torch.equal(o.edge_attr, torch.cat([ea[1:], torch.ones(3, 2)]))
    assert o.pos is None and torch.equal(o.y, torch.tensor([1.0]))


def test_pre_transform_batched_equals_per_graph_calls():
    """N1: the batched dataset pre_transform returns the same per-graph Data as calling create_subgraphs on each graph."""
    from esc_gnn_b200 import synth
    from esc_gnn_b200.data import Data
    from esc_gnn_b200.dataset import pre_transform_batched
    from esc_gnn_b200.transform import create_subgraphs
    graphs = []
    for i in range(37):
        g = synth.make_graph(4 if i % 2 else 1, 900 + i)
        graphs.append(Data(x=torch.as_tensor(g['x']), edge_index=torch.as_tensor(g['edge_index']),
                           edge_attr=torch.as_tensor(g['edge_attr']).float() if 'edge_attr' in g else None,
                           y=torch.as_tensor(g['y'])))
    for sl, rd in ((True, True), (False, False)):
        got = pre_transform_batched(graphs, h=3, use_rd=rd, self_loop=sl, chunk=16)
        for d, o in zip(graphs, got):
            w = create_subgraphs(d, 3, use_rd=rd, self_loop=sl)
            for k in ('edge_index', 'pos_enc', 'pos_index', 'pos_batch'):
                assert torch.equal(o[k], w[k]), k
            assert (o.edge_attr is None) == (w.edge_attr is None) and (o.edge_attr is None or torch.equal(o.edge_attr, w.edge_attr))


def test_pre_transform_batched_single_edge_graphs_inside_a_chunk():
    """A graph with exactly ONE output edge (one node + self_loop, or one directed edge) has last local ordinal 0, the same as
    the next graph's first: record boundaries must come from the batch-wide edge ordinal, not from ordinal drops."""
    from esc_gnn_b200 import synth
    from esc_gnn_b200.data import Data
    from esc_gnn_b200.dataset import pre_transform_batched
    from esc_gnn_b200.transform import create_subgraphs
    one_node = Data(x=torch.ones(1, 1), edge_index=torch.zeros((2, 0), dtype=torch.long), y=torch.tensor([0.0]))
    one_node.num_nodes = 1
    one_edge = Data(x=torch.ones(2, 1), edge_index=torch.tensor([[0], [1]]), y=torch.tensor([1.0]))

    def big(i):
        g = synth.make_graph(2, 40 + i)
        return Data(x=torch.as_tensor(g['x']), edge_index=torch.as_tensor(g['edge_index']), y=torch.as_tensor(g['y']).view(1))
    # self_loop=True: the one-node graph yields the single edge (0,0); self_loop=False: only the one-edge graph qualifies
    for sl, graphs in ((True, [big(0), one_node, one_node, big(1), one_edge, one_node, big(2)]),
                       (False, [one_edge, big(0), one_edge, one_edge, big(1), one_edge])):
        got = pre_transform_batched(graphs, h=2, use_rd=False, self_loop=sl, chunk=64)
        assert len(got) == len(graphs)
        for d, o in zip(graphs, got):
            w = create_subgraphs(d, 2, use_rd=False, self_loop=sl)
            for k in ('edge_index', 'pos_enc', 'pos_index', 'pos_batch'):
                assert torch.equal(o[k], w[k]), (sl, k)


def test_encoded_dataset_processes_once_and_serves_from_cache(tmp_path):
    """N1: EncodedDataset.process() (batched encoder -> collate -> torch.save) then a second open from the cache alone;
    every served graph equals the oracle's per-graph encoding, and the loader batches it like any Data list."""
    from esc_gnn_b200 import synth
    from esc_gnn_b200.data import Data
    from esc_gnn_b200.dataloader import DataLoader
    from esc_gnn_b200.dataset import EncodedDataset
    from tests import model_util as MU
    raw = []
    for i in range(300, 323):
        g = synth.make_graph(2, i)
        raw.append(Data(x=torch.as_tensor(g['x']), edge_index=torch.as_tensor(g['edge_index']),
                        edge_attr=torch.as_tensor(g['edge_attr']), y=torch.as_tensor(g['y']).view(1)))
    ds = EncodedDataset(str(tmp_path), raw, h=3, use_rd=True, self_loop=False, chunk=10, name='zincish')

    def boom():
        raise AssertionError('cache ignored')
    ds2 = EncodedDataset(str(tmp_path), boom, h=3, use_rd=True, self_loop=False, name='zincish')
    assert len(ds) == len(ds2) == 23
    want = MU.graph_dicts(2, 300, 23)
    for i in (0, 7, 22):
        for k in ('edge_index', 'pos_enc', 'pos_index', 'pos_batch', 'x', 'edge_attr'):
            assert torch.equal(ds2[i][k], want[i][k]), (i, k)
        assert ds2[i].num_nodes == want[i]['num_nodes']
    sub = ds2.shuffle(torch.Generator().manual_seed(0))[:8]
    assert len(sub) == 8
    b = next(iter(DataLoader(list(sub), batch_size=8)))
    assert b.num_graphs == 8 and int(b.pos_batch[-1]) + 1 == b.edge_index.size(1)


def test_all_pairs_spd_matches_fixture_and_oracle():
    """N4: `attn_bias` (GraphGPS/graphgps/loader/utils_escgnn.py:29-38) -- bit-exact against the networkx fixture, and
    against the oracle on a batch of config-5 graphs (up to 500 nodes) through the batched entry point."""
    import os
    from esc_gnn_b200 import synth
    from esc_gnn_b200.transform import all_pairs_spd, all_pairs_spd_batch
    from oracle import encode_ref
    fix = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'spd.npz'))
    for nm in sorted({k.split('/')[0] for k in fix.files}):
        got = all_pairs_spd(torch.as_tensor(fix[nm + '/edge_index']), int(fix[nm + '/n'][0]))
        assert got.dtype == torch.int64 and np.array_equal(got.numpy(), fix[nm + '/attn_bias']), nm
    src, dst, eptr, nptr = synth.make_batch_arrays(5, 40, 12)
    flat, optr = all_pairs_spd_batch(src, dst, eptr, nptr)
    flat = flat.cpu().numpy()
    for g in range(12):
        ei = np.stack([src[eptr[g]:eptr[g + 1]], dst[eptr[g]:eptr[g + 1]]])
        want = encode_ref.all_pairs_spd(ei, int(nptr[g + 1] - nptr[g]))
        assert np.array_equal(flat[optr[g]:optr[g + 1]], want), g
    with pytest.raises(RuntimeError):
        all_pairs_spd(torch.tensor([[0, 5], [1, 0]]), 3)          # node id outside [0, n)


def test_create_subgraphs_gps_adds_attn_bias():
    from esc_gnn_b200 import synth
    from esc_gnn_b200.data import Data
    from esc_gnn_b200.transform import create_subgraphs, create_subgraphs_gps
    from oracle import encode_ref
    g = synth.make_graph(2, 77)
    d = Data(x=torch.as_tensor(g['x']), edge_index=torch.as_tensor(g['edge_index']), edge_attr=torch.as_tensor(g['edge_attr']),
             y=torch.as_tensor(g['y']).view(1))
    a, b = create_subgraphs(d, 2, use_rd=True, self_loop=True), create_subgraphs_gps(d, 2, use_rd=True, self_loop=True)
    for k in ('edge_index', 'pos_enc', 'pos_index', 'pos_batch', 'edge_attr'):
        assert torch.equal(a[k], b[k])
    assert np.array_equal(b.attn_bias.numpy(), encode_ref.all_pairs_spd(g['edge_index'], g['num_nodes']))


@pytest.mark.parametrize('config', [1, 2])
def test_pipelined_host_encoder_matches_oracle(config):
    """HostEncoder.stream: chunks submitted on alternating slots (D2H of chunk k under the kernels of chunk k+1), compact
    records fetched into pinned arenas, the int64 triple expanded on the host -- against the C oracle, for chunks of different
    sizes (arena growth between calls) and both ordinal conventions."""
    from esc_gnn_b200 import synth
    from esc_gnn_b200.transform import HostEncoder
    fl = synth.ENCODER_FLAGS[config]
    sizes = [40, 130, 7, 64, 1]
    chunks, start = [], 6000
    for n in sizes:
        chunks.append(synth.make_batch_arrays(config, start, n))
        start += n
    for local in (False, True):
        enc = HostEncoder(fl['h'], fl['use_rd'], fl['self_loop'], local_ordinals=local)
        seen = 0
        for (src, dst, eptr, nptr), r in zip(chunks, enc.stream(iter(chunks))):
            want = _oracle_batch(src, dst, eptr, nptr, fl['h'], fl['use_rd'], fl['self_loop'])
            assert np.array_equal(r.edge_index.numpy(), want[0])
            assert np.array_equal(r.pos_enc.numpy(), want[1]) and np.array_equal(r.pos_index.numpy(), want[2])
            if not local:
                assert np.array_equal(r.pos_batch.numpy(), want[3])
            else:       # per-graph ordinals: what a per-graph Data object holds before collation
                ep = r.edge_ptr.numpy()
                g_of_e = np.searchsorted(ep, want[3], side='right') - 1
                assert np.array_equal(r.pos_batch.numpy(), want[3] - ep[g_of_e])
            assert r.nnz == want[1].shape[0] and int(r.rec_nnz.sum()) == r.nnz
            seen += 1
        assert seen == len(chunks)


@pytest.mark.parametrize('config', [2, 4, 1])
def test_rd_pendant_tree_peeling_leaves_the_histograms_unchanged(config):
    """ego_rd peels pendant trees before the linear algebra (R(r, leaf) = R(r, parent) + 1): same records as the full solve of
    every pair system, on molecule-shaped batches and on a multigraph whose doubled edges must NOT be peeled (weight 2)."""
    from esc_gnn_b200 import _lib, synth
    from esc_gnn_b200.transform import encode_batch
    L = _lib.lib()
    fl = synth.ENCODER_FLAGS[config]
    src, dst, eptr, nptr = synth.make_batch_arrays(config, 8000, 192)
    # append one hand-made multigraph: a triangle 0-1-2 with a pendant path 2-3-4 and a DOUBLE pendant edge 0=5 (both directions twice)
    und = [(0, 1), (1, 2), (0, 2), (2, 3), (3, 4), (0, 5), (0, 5)]
    ms = np.array([a for a, b in und] + [b for a, b in und], dtype=np.int64)
    md = np.array([b for a, b in und] + [a for a, b in und], dtype=np.int64)
    src, dst = np.concatenate([src, ms]), np.concatenate([dst, md])
    eptr, nptr = np.append(eptr, eptr[-1] + len(ms)), np.append(nptr, nptr[-1] + 6)
    args = (torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda(), torch.as_tensor(eptr), torch.as_tensor(nptr), fl['h'], True,
            fl['self_loop'])
    was = L.escgnn_set_rd_peel(1)
    try:
        a = encode_batch(*args)
        L.escgnn_set_rd_peel(0)
        b = encode_batch(*args)
    finally:
        L.escgnn_set_rd_peel(was)
    for k in ('pos_enc', 'pos_index', 'pos_batch'):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    want = _oracle_batch(src, dst, eptr, nptr, fl['h'], True, fl['self_loop'])
    assert np.array_equal(a.pos_enc.cpu().numpy(), want[1]) and np.array_equal(a.pos_index.cpu().numpy(), want[2])


@pytest.mark.parametrize('config,count', [(2, 300), (4, 200), (1, 64), (6, 100)])
def test_rd_cycle_space_fast_path_equals_the_general_solver(config, count):
    """ego_rd_fast_kernel (csrc/rd_fast.cuh: a thread per pair system, resistance distances from the cycle space) in front of the
    LDL^T / Takahashi solver: same records with the fast path on and off, and equal to the oracle -- on molecule-shaped batches
    (everything solved by the fast path), on count-shaped ones (mostly declined: > 4 cycles per ego-net), with 42..130-node graphs
    (second size class), a multigraph and an asymmetric-free single-node graph in the batch."""
    from esc_gnn_b200 import _lib, synth
    from esc_gnn_b200.transform import encode_batch
    L = _lib.lib()
    fl = synth.ENCODER_FLAGS[config]
    src, dst, eptr, nptr = synth.make_batch_arrays(config, 12000, count)
    rng = np.random.Generator(np.random.PCG64(3))
    extra = []
    for n in (60, 128, 130, 42):                     # 130 nodes: beyond the fast path's 128
        extra.append((synth.symmetrise(synth.random_graph(rng, n, n + 2, max_degree=4)), n))
    und = [(0, 1), (1, 2), (0, 2), (2, 3), (3, 4), (0, 5), (0, 5)]        # doubled edge 0=5: multigraph -> general solver
    extra.append((np.array([[a for a, b in und] + [b for a, b in und], [b for a, b in und] + [a for a, b in und]], dtype=np.int64), 6))
    extra.append((np.array([[0, 1], [1, 0]], dtype=np.int64), 2))
    for ei, n in extra:
        src, dst = np.concatenate([src, ei[0]]), np.concatenate([dst, ei[1]])
        eptr, nptr = np.append(eptr, eptr[-1] + ei.shape[1]), np.append(nptr, nptr[-1] + n)
    args = (torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda(), torch.as_tensor(eptr), torch.as_tensor(nptr), fl['h'], True,
            fl['self_loop'])
    was = L.escgnn_set_rd_fast(1)
    try:
        a = encode_batch(*args)
        L.escgnn_set_rd_fast(0)
        b = encode_batch(*args)
    finally:
        L.escgnn_set_rd_fast(was)
    for k in ('pos_enc', 'pos_index', 'pos_batch'):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    want = _oracle_batch(src, dst, eptr, nptr, fl['h'], True, fl['self_loop'])
    assert np.array_equal(a.pos_enc.cpu().numpy(), want[1]) and np.array_equal(a.pos_index.cpu().numpy(), want[2])
