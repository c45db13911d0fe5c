"""Deterministic synthetic graphs of the BASELINE.json config shapes.

The reference's dataset blobs for configs 1-3 are absent, so every measurement and parity case
uses these generators.  Graph `i` of config `c` is drawn from `numpy.random.Generator(PCG64(1000*c + i))`.

G(n, m): uniform random recursive spanning tree (parent(i) ~ U{0..i-1}) plus m-(n-1) distinct random chords,
no loops / multi-edges, optional max-degree cap, symmetrised `edge_index` (forward block, then reversed block).
Shapes follow the callers of the reference transform: `GraphCountDataset.py:69-84` (x = ones[n,10]),
`dataset_zinc.py:64-88` (integer atom / bond types), `run_ogb_mol.py:319-332` (9 atom / 3 bond columns).
"""
import numpy as np

ATOM_DIMS = (119, 4, 12, 12, 10, 6, 6, 2, 2)   # ogb get_atom_feature_dims() is third-party; parameterised here
BOND_DIMS = (5, 6, 2)

# config id -> (h, use_rd, self_loop) exactly as the reference scripts call create_subgraphs
ENCODER_FLAGS = {
    1: dict(h=3, use_rd=True, self_loop=True),     # run_graphcount.py:404-408
    2: dict(h=3, use_rd=True, self_loop=False),    # run_zinc.py:141-146 (use_rd always on, :42)
    3: dict(h=4, use_rd=True, self_loop=True),     # count_graphlet
    4: dict(h=4, use_rd=True, self_loop=True),     # run_ogb_mol.py:329-332
    5: dict(h=3, use_rd=False, self_loop=False),   # sweep: h in 1..4, rd off
    6: dict(h=3, use_rd=True, self_loop=True),     # run_qm9.py:203-205 (QM9 variant)
    7: dict(h=3, use_rd=True, self_loop=True),     # run_csl.py:79-80 (graph classification, kernel/gin.py variant, N3)
}
KGIN_FEATURES, KGIN_CLASSES = 7, 3                 # graph-classification variant (kernel/gin.py:200-379): dataset.num_features / num_classes
QM9_FEATURES = 11                                  # dataset.num_features of the reference's QM9 (run_qm9.py:216-231)


def random_graph(rng, n, m, max_degree=199):
    """Undirected simple graph with n nodes and (up to) m edges; returns int64 array [m', 2] with a < b per row."""
    n = int(n)
    if n <= 1:
        return np.zeros((0, 2), dtype=np.int64)
    m = int(min(m, n * (n - 1) // 2))
    deg = np.zeros(n, dtype=np.int64)
    if max_degree >= n:  # cap cannot bind: fully vectorised tree
        par = (rng.random(n - 1) * np.arange(1, n)).astype(np.int64)
        np.add.at(deg, par, 1)
        deg[1:] += 1
    else:
        par = np.empty(n - 1, dtype=np.int64)
        for i in range(1, n):
            p = int(rng.integers(0, i))
            tries = 0
            while deg[p] >= max_degree and tries < 64:
                p = int(rng.integers(0, i)); tries += 1
            if deg[p] >= max_degree:
                p = int(np.argmin(deg[:i]))
            par[i - 1] = p
            deg[p] += 1; deg[i] += 1
    child = np.arange(1, n, dtype=np.int64)
    keys = set((np.minimum(par, child) * n + np.maximum(par, child)).tolist())
    need = m - (n - 1)
    guard = 0
    while need > 0 and guard < 50:
        guard += 1
        a = rng.integers(0, n, size=2 * need + 8)
        b = rng.integers(0, n, size=2 * need + 8)
        for x, y in zip(a.tolist(), b.tolist()):
            if need == 0:
                break
            if x == y or deg[x] >= max_degree or deg[y] >= max_degree:
                continue
            k = min(x, y) * n + max(x, y)
            if k in keys:
                continue
            keys.add(k); deg[x] += 1; deg[y] += 1; need -= 1
    keys = np.array(sorted(keys), dtype=np.int64)
    return np.stack([keys // n, keys % n], axis=1)


def symmetrise(und):
    """[m,2] undirected pairs -> edge_index [2, 2m]: forward block then reversed block."""
    if und.shape[0] == 0:
        return np.zeros((2, 0), dtype=np.int64)
    return np.concatenate([und.T, und.T[::-1]], axis=1).astype(np.int64)


def make_graph(config, i):
    """One synthetic graph of `config` (1..7) as a dict of numpy arrays: edge_index, num_nodes, x, y[, edge_attr]."""
    rng = np.random.Generator(np.random.PCG64(1000 * config + i))
    if config in (1, 3):
        n = int(rng.integers(10, 31))
        und = random_graph(rng, n, int(round(1.66 * n)))
        ei = symmetrise(und)
        return dict(edge_index=ei, num_nodes=n, x=np.ones((n, 10), dtype=np.float32),
                    y=rng.random(n).astype(np.float32))
    if config == 2:
        n = int(np.clip(round(rng.normal(23.2, 4.5)), 9, 37))
        und = random_graph(rng, n, n - 1 + int(rng.integers(0, 4)), max_degree=4)
        ei = symmetrise(und)
        ea = rng.integers(1, 4, size=und.shape[0]).astype(np.int64)
        return dict(edge_index=ei, num_nodes=n, x=rng.integers(0, 28, size=n).astype(np.int64),
                    edge_attr=np.concatenate([ea, ea]), y=np.float32(rng.normal()))
    if config == 4:
        n = int(np.clip(round(rng.lognormal(3.15, 0.35)), 6, 120))
        und = random_graph(rng, n, n - 1 + int(rng.integers(0, 5)), max_degree=4)
        ei = symmetrise(und)
        x = np.stack([rng.integers(0, d, size=n) for d in ATOM_DIMS], axis=1).astype(np.int64)
        ea = np.stack([rng.integers(0, d, size=und.shape[0]) for d in BOND_DIMS], axis=1).astype(np.int64)
        return dict(edge_index=ei, num_nodes=n, x=x, edge_attr=np.concatenate([ea, ea], axis=0),
                    y=np.float32(rng.random() < 0.03))
    if config == 6:      # QM9-shaped: small 3-D molecules, continuous node features, one-hot bond types, per-graph target
        n = int(rng.integers(4, 30))
        und = random_graph(rng, n, n - 1 + int(rng.integers(0, 3)), max_degree=4)
        bond = np.eye(4, dtype=np.float32)[rng.integers(0, 4, size=und.shape[0])]
        return dict(edge_index=symmetrise(und), num_nodes=n, x=rng.random((n, QM9_FEATURES)).astype(np.float32),
                    pos=(1.5 * rng.normal(size=(n, 3))).astype(np.float32), node_type=rng.integers(0, 5, size=n).astype(np.int64),
                    edge_attr=np.concatenate([bond, bond], axis=0), y=np.float32(rng.normal()))
    if config == 7:      # graph classification (CSL / EXP / TU style): continuous node features, one integer class per graph
        n = int(rng.integers(10, 31))
        und = random_graph(rng, n, int(round(1.66 * n)))
        return dict(edge_index=symmetrise(und), num_nodes=n, x=rng.random((n, KGIN_FEATURES)).astype(np.float32),
                    y=np.int64(rng.integers(0, KGIN_CLASSES)))
    if config == 5:
        n = int(rng.integers(25, 501))
        und = random_graph(rng, n, int(1.25 * n))
        return dict(edge_index=symmetrise(und), num_nodes=n)
    raise ValueError('unknown config %r' % (config, ))


def make_batch_arrays(config, start, count):
    """`count` graphs of `config` packed for the batched encoder: (src, dst, edge_ptr, node_ptr) int64,
    node ids graph-local.  This is the raw-input side of the encoder hot path."""
    srcs, dsts, eptr, nptr = [], [], [0], [0]
    for i in range(start, start + count):
        g = make_graph(config, i)
        srcs.append(g['edge_index'][0]); dsts.append(g['edge_index'][1])
        eptr.append(eptr[-1] + g['edge_index'].shape[1]); nptr.append(nptr[-1] + g['num_nodes'])
    return (np.concatenate(srcs), np.concatenate(dsts), np.asarray(eptr, dtype=np.int64),
            np.asarray(nptr, dtype=np.int64))
