"""Generate golden fixtures by executing the UNMODIFIED reference encoder from /root/reference.

Run in the build container only (needs the read-only reference mount):
    python tests/golden/make_golden.py
It imports `/root/reference/utils_edge_efficient.py` as-is under the PyG stand-in in `tests/_pyg_shim`
(torch_geometric / torch_scatter are not installable here) and freezes
`create_subgraphs(...)` outputs in a compact layout:

    <case>/ei   int32 [2, E_in]   input edge_index          <case>/meta int64 [n, h, use_rd, self_loop]
    <case>/eo   int32 [2, E_out]  output edge_index (E1)    <case>/nnz  int32 [E_out]  entries per edge
    <case>/idx  int16 [nnz]       pos_index                 <case>/cnt  int32 [nnz]    pos_enc
(pos_batch is `repeat(arange(E_out), nnz)`; the generator asserts that.)

With use_rd the [400,500) block is the literal float32-LAPACK output of the reference (not reproducible,
SURVEY F7); tests compare it under parity policy E5, every other index bit-exactly.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'tests', '_pyg_shim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, ROOT)

from torch_geometric.data import Data  # noqa: E402  (the stand-in)
import utils_edge_efficient as REF  # noqa: E402  (the unmodified reference)

from esc_gnn_b200 import synth  # noqa: E402


def run_reference(ei, n, h, use_rd, self_loop):
    d = Data(x=torch.ones(n, 1), edge_index=torch.as_tensor(np.asarray(ei), dtype=torch.long))
    d.num_nodes = n
    o = REF.create_subgraphs(d, h, use_rd=use_rd, self_loop=self_loop)
    eo = o.edge_index.numpy()
    pb = o.pos_batch.numpy()
    nnz = np.bincount(pb, minlength=eo.shape[1]).astype(np.int32)
    assert np.array_equal(pb, np.repeat(np.arange(eo.shape[1]), nnz))
    return eo.astype(np.int32), nnz, o.pos_index.numpy().astype(np.int16), o.pos_enc.numpy().astype(np.int32)


def add_case(store, name, ei, n, h, use_rd, self_loop):
    eo, nnz, idx, cnt = run_reference(ei, n, h, use_rd, self_loop)
    store[name + '/ei'] = np.asarray(ei, dtype=np.int32).reshape(2, -1)
    store[name + '/meta'] = np.array([n, h, int(use_rd), int(self_loop)], dtype=np.int64)
    store[name + '/eo'] = eo
    store[name + '/nnz'] = nnz
    store[name + '/idx'] = idx
    store[name + '/cnt'] = cnt


def save(fname, store):
    path = os.path.join(HERE, fname)
    np.savez_compressed(path, **store)
    print('%-12s %4d cases %8.1f KB' % (fname, len(store) // 6, os.path.getsize(path) / 1024))


def kat1():
    und = [(0, 1), (1, 2), (2, 3), (1, 4), (2, 4)]
    ei = np.array(und + [(b, a) for a, b in und]).T
    s = {}
    for sl in (0, 1):
        for rd in (0, 1):
            for h in (1, 2, 3, 4):
                add_case(s, 'kat1_h%d_rd%d_sl%d' % (h, rd, sl), ei, 5, h, bool(rd), bool(sl))
    save('kat1.npz', s)


def edge_cases():
    s = {}
    tri = np.array([[0, 1, 1, 2, 2, 0], [1, 0, 2, 1, 0, 2]])
    path_dir = np.array([[0, 1, 2, 3], [1, 2, 3, 4]])                      # directed path: BFS walks target->source (F10)
    dup = np.array([[0, 1, 0, 1, 1, 2, 2, 1], [1, 0, 1, 0, 2, 1, 1, 2]])   # duplicate edges count twice
    loops = np.array([[0, 0, 1, 1, 2, 2], [0, 1, 0, 1, 2, 0]])             # pre-existing loops, asymmetric
    iso = np.array([[0, 1], [1, 0]])                                       # nodes 2..4 isolated
    star = np.array([[0] * 7 + list(range(1, 8)), list(range(1, 8)) + [0] * 7])
    for h in (1, 2, 3, 4):
        for sl in (0, 1):
            add_case(s, 'tri_h%d_sl%d' % (h, sl), tri, 3, h, False, bool(sl))
            add_case(s, 'dirpath_h%d_sl%d' % (h, sl), path_dir, 5, h, False, bool(sl))
            add_case(s, 'dup_h%d_sl%d' % (h, sl), dup, 3, h, False, bool(sl))
            add_case(s, 'loops_h%d_sl%d' % (h, sl), loops, 3, h, False, bool(sl))
            add_case(s, 'iso_h%d_sl%d' % (h, sl), iso, 5, h, False, bool(sl))
            add_case(s, 'star_h%d_sl%d' % (h, sl), star, 8, h, False, bool(sl))
        add_case(s, 'empty_h%d_sl1' % h, np.zeros((2, 0), dtype=np.int64), 4, h, False, True)
        add_case(s, 'single_h%d_sl1' % h, np.zeros((2, 0), dtype=np.int64), 1, h, False, True)
        for sl in (0, 1):                                                  # rd on symmetric small graphs
            add_case(s, 'tri_rd_h%d_sl%d' % (h, sl), tri, 3, h, True, bool(sl))
            add_case(s, 'star_rd_h%d_sl%d' % (h, sl), star, 8, h, True, bool(sl))
            add_case(s, 'duprd_h%d_sl%d' % (h, sl), dup, 3, h, True, bool(sl))
    save('edge_cases.npz', s)


def sr25():
    """KAT-2: the 15 strongly regular graphs the reference ships (data/sr25/raw/sr251256.g6; run_sr.py:76-78)."""
    import networkx as nx
    graphs = nx.read_graph6('/root/reference/data/sr25/raw/sr251256.g6')
    s = {}
    for gi, g in enumerate(graphs):
        a = nx.to_numpy_array(g)
        ei = np.array(np.nonzero(a))            # SRDataset.py:30-47 builds edge_index from the adjacency matrix
        for h in (1, 2, 3):
            add_case(s, 'sr25_g%02d_h%d' % (gi, h), ei, 25, h, False, True)
    save('sr25.npz', s)


def configs():
    counts = {1: 24, 2: 40, 3: 10, 4: 16}
    for c, k in counts.items():
        s = {}
        fl = synth.ENCODER_FLAGS[c]
        for i in range(k):
            g = synth.make_graph(c, i)
            add_case(s, 'cfg%d_g%03d' % (c, i), g['edge_index'], g['num_nodes'], fl['h'], fl['use_rd'], fl['self_loop'])
        save('cfg%d.npz' % c, s)
    s = {}
    picked = [i for i in range(60) if synth.make_graph(5, i)['num_nodes'] <= 260][:8]
    for i in picked:
        g = synth.make_graph(5, i)
        for h in (1, 2, 3, 4):
            sl = (i + h) % 2
            add_case(s, 'cfg5_g%03d_h%d_sl%d' % (i, h, sl), g['edge_index'], g['num_nodes'], h, False, bool(sl))
    save('cfg5.npz', s)


if __name__ == '__main__':
    kat1()
    edge_cases()
    sr25()
    configs()
