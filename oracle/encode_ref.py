"""ORACLE (test infrastructure, NOT product code) -- numpy restatement of the ESC-GNN structural encoder.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module; the product package `esc_gnn_b200` never does (it fails loudly when its CUDA library is missing).

What it restates (all line numbers are /root/reference/utils_edge_efficient.py):
  * E1  self-loop rewrite .......................... :33-38   (`remove_self_loops` + `add_self_loops`)
  * E2  bounded BFS with hop labels ................ :201-294 (`k_hop_subgraph`, node_label='hop';
                                                     walks target->source: `col, row = edge_index`, :207-210)
  * E3  union of the two balls, (d0, d1) labels .... :52-67   (phantom duplicate root when u == v)
  * E4  degree / distance / distance-pair histograms  :86, :122-144
  * E5  resistance-distance block .................. :92-107, :130-131
  * E6  output assembly ............................. :146-152

Parity status: PINNED.  The reference ships no tests for this path, so the pins are outputs of
the reference's own unmodified source, executed in the build container under the PyG stand-in of
`tests/_pyg_shim` by `tests/golden/make_golden.py` and frozen under `tests/golden/*.npz`
(KAT-1, SR25 KAT-2, random graphs of every config shape).  Integer blocks: bit-exact against those fixtures.
rd block: the reference computes it with a float32 LAPACK pinv and truncates, which is not reproducible
(SURVEY F7); `rd_mode='fp64'` (the parity policy E5) computes the same formula in float64 and bins
`trunc(float32(rd))`; tests assert that every disagreement with the literal reference sits on an
integer-valued resistance distance.  `rd_mode='ref32'` re-enacts the float32 path for diagnostics.
"""
import numpy as np

DEG_BINS = 200      # F.one_hot(sub_degree, 200)            :128
DIST_BINS = 100     # F.one_hot(z, 100) -> two blocks       :128
RD_BINS = 100       # F.one_hot(rd, 100)                    :131
CODE_BINS = 1300    # F.one_hot(code, 1300)                 :138
CODE_W = (216, 36, 6, 1)


def rewrite_self_loops(edge_index, num_nodes):
    """E1 (:33-38): drop existing loops, then append (i, i) for every node, in node order."""
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    keep = ei[0] != ei[1]
    loops = np.arange(num_nodes, dtype=np.int64)
    return np.concatenate([ei[:, keep], np.stack([loops, loops])], axis=1), keep


def _in_neighbours(ei, num_nodes):
    """BFS neighbours of t are the sources s of edges (s, t) (:207-210, :222-226)."""
    nbr = [[] for _ in range(num_nodes)]
    for s, t in zip(ei[0].tolist(), ei[1].tolist()):
        nbr[t].append(s)
    return nbr


def bfs_hops(root, h, nbr):
    """E2 (:217-243): hop distance root->w for every w within h hops; returns dict node -> hop."""
    dist = {root: 0}
    frontier = [root]
    for level in range(h):
        nxt = []
        for t in frontier:
            for s in nbr[t]:
                if s not in dist:
                    dist[s] = level + 1
                    nxt.append(s)
        if not nxt:
            break
        frontier = nxt
    return dist


def _resistance_bins(n_sub, sub_src, sub_dst, rd_mode):
    """E5 (:92-107, :130-131). Returns int64 bin per sub-node (index 0 is the first root)."""
    dt = np.float32 if rd_mode == 'ref32' else np.float64
    adj = np.zeros((n_sub, n_sub), dtype=dt)
    np.add.at(adj, (sub_src, sub_dst), 1)          # duplicates add up (coo -> csr), :94-96
    # scipy.sparse.csgraph.laplacian (default: in-degree, loops ignored): L = diag(colsum - diag) - A
    w = adj.sum(axis=0) - np.diag(adj)
    lap = -adj
    np.fill_diagonal(lap, w)
    if rd_mode == 'ref32':
        from scipy import linalg
        linv = linalg.pinv(lap)
    else:
        linv = np.linalg.pinv(lap, hermitian=bool(np.array_equal(lap, lap.T)))
    rd = linv[0, 0] + np.diag(linv) - linv[0, :] - linv[:, 0]
    rd32 = np.asarray(rd, dtype=np.float32)        # torch.FloatTensor(...), :105
    return np.trunc(rd32).astype(np.int64), rd     # .long(), :131


def encode_graph(edge_index, num_nodes, h, use_rd=False, self_loop=False, rd_mode='fp64', rd_out=None):
    """Encode one graph. Returns (edge_index_out[2,E], pos_enc, pos_index, pos_batch) as int64 arrays.

    Mirrors `create_subgraphs(data, h, use_rd=..., self_loop=...)` for the fields E6 produces.
    `rd_out`: optional list; receives one float array of raw resistance distances per edge (E5 policy checks).
    Raises ValueError where the reference's `F.one_hot` would raise (degree >= 200, h >= 5, rd >= 100).
    """
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    if self_loop:
        ei, _ = rewrite_self_loops(ei, num_nodes)
    E = ei.shape[1]
    if 216 * (h + 1) + 36 * (h + 1) + 6 * (h + 1) + (h + 1) >= CODE_BINS:
        raise ValueError('h too large for the 1300-bin distance-pair code (reference one_hot would raise)')
    src, dst = ei[0].tolist(), ei[1].tolist()
    nbr = _in_neighbours(ei, num_nodes)
    balls = {}
    off = (DEG_BINS + 2 * DIST_BINS + RD_BINS) if use_rd else (DEG_BINS + 2 * DIST_BINS)
    width = off + CODE_BINS
    pos_enc, pos_index, pos_batch = [], [], []
    for e in range(E):
        u, v = src[e], dst[e]
        for r in (u, v):
            if r not in balls:
                balls[r] = bfs_hops(r, h, nbr)
        bu, bv = balls[u], balls[v]
        # E3: node list [u, v] ++ (B_u minus seen) ++ (B_v minus seen); u == v gives a phantom duplicate (F8)
        nodes = [u, v]
        seen = {u, v}
        for w in bu:
            if w not in seen:
                nodes.append(w); seen.add(w)
        for w in bv:
            if w not in seen:
                nodes.append(w); seen.add(w)
        z = np.array([[bu.get(w, h + 1), bv.get(w, h + 1)] for w in nodes], dtype=np.int64)
        relabel = {}
        for i, w in enumerate(nodes):
            relabel[w] = i                           # last write wins: u == v maps to slot 1
        in_u = np.zeros(num_nodes, dtype=bool); in_u[list(bu)] = True
        in_v = np.zeros(num_nodes, dtype=bool); in_v[list(bv)] = True
        # F = induced(B_u) OR induced(B_v) over the edge LIST (:55, :283-285)
        fmask = (in_u[ei[0]] & in_u[ei[1]]) | (in_v[ei[0]] & in_v[ei[1]])
        fs = np.array([relabel[a] for a in ei[0][fmask].tolist()], dtype=np.int64)
        fd = np.array([relabel[b] for b in ei[1][fmask].tolist()], dtype=np.int64)
        n_sub = len(nodes)
        enc = np.zeros(width, dtype=np.int64)
        deg = np.bincount(fs, minlength=n_sub)       # degree(edge_index_[0]), :86
        if deg.size and deg.max() >= DEG_BINS:
            raise ValueError('sub_degree >= 200 (reference one_hot would raise)')
        np.add.at(enc, deg, 1)
        np.add.at(enc, DEG_BINS + z[:, 0], 1)
        np.add.at(enc, DEG_BINS + DIST_BINS + z[:, 1], 1)
        if use_rd:
            rb, rd_raw = _resistance_bins(n_sub, fs, fd, rd_mode)
            if rd_out is not None:
                rd_out.append(rd_raw)
            if rb.min() < 0 or rb.max() >= RD_BINS:
                raise ValueError('rd bin outside [0, 100) (reference one_hot would raise)')
            np.add.at(enc, DEG_BINS + 2 * DIST_BINS + rb, 1)
        nl = fs != fd                                # remove_self_loops(sg.edge_index), :138
        code = (CODE_W[0] * z[fs[nl], 0] + CODE_W[1] * z[fs[nl], 1]
                + CODE_W[2] * z[fd[nl], 0] + CODE_W[3] * z[fd[nl], 1])
        np.add.at(enc, off + code, 1)
        idx = np.nonzero(enc)[0]
        pos_index.append(idx)
        pos_enc.append(enc[idx])
        pos_batch.append(np.full(idx.size, e, dtype=np.int64))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)
    return ei, cat(pos_enc), cat(pos_index), cat(pos_batch)


def all_pairs_spd(edge_index, num_nodes, unreachable=100):
    """`attn_bias` of the GraphGPS twin (GraphGPS/graphgps/loader/utils_escgnn.py:29-38): shortest-path length between
    every pair of nodes of the UNDIRECTED graph (`to_networkx(data, to_undirected=True)` then networkx
    `all_pairs_shortest_path_length`, a third-party routine: plain BFS per source), `unreachable` (100) where no path
    exists, flattened row-major to int64 [n*n].  Pinned against networkx itself by tests/golden/make_golden_spd.py."""
    n = int(num_nodes)
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    nbr = [set() for _ in range(n)]
    for s, t in zip(ei[0].tolist(), ei[1].tolist()):
        if s != t:
            nbr[s].add(t); nbr[t].add(s)
    out = np.full((n, n), unreachable, dtype=np.int64)
    for r in range(n):
        out[r, r] = 0
        frontier, d = [r], 0
        seen = {r}
        while frontier:
            d += 1
            nxt = []
            for w in frontier:
                for s in nbr[w]:
                    if s not in seen:
                        seen.add(s); out[r, s] = d; nxt.append(s)
            frontier = nxt
    return out.reshape(-1)
