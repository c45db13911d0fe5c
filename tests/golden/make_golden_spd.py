"""Golden `attn_bias` vectors from networkx -- the routine the reference itself calls
(/root/reference/GraphGPS/graphgps/loader/utils_escgnn.py:29-38; `to_networkx(..., to_undirected=True)` is restated as
"one undirected edge per directed pair").  Run in the build container:  python tests/golden/make_golden_spd.py"""
import os
import sys

import networkx as nx
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from esc_gnn_b200 import synth  # noqa: E402


def cases():
    out = {}
    for cfg, i in ((1, 3), (2, 5), (4, 7), (5, 11), (6, 2)):
        g = synth.make_graph(cfg, i)
        out['cfg%d_%d' % (cfg, i)] = (g['edge_index'], g['num_nodes'])
    # disconnected: two components + isolated nodes, a self-loop, and a one-directional edge
    out['disconnected'] = (np.array([[0, 1, 1, 2, 4, 5, 3, 6], [1, 0, 2, 1, 5, 4, 3, 7]], dtype=np.int64), 10)
    out['path40'] = (np.stack([np.arange(39), np.arange(1, 40)]).astype(np.int64), 40)
    out['single'] = (np.zeros((2, 0), dtype=np.int64), 1)
    return out


def reference_attn_bias(edge_index, n):
    G = nx.Graph()
    G.add_nodes_from(range(n))
    G.add_edges_from(zip(edge_index[0].tolist(), edge_index[1].tolist()))
    SPT = dict(nx.all_pairs_shortest_path_length(G))
    a = np.zeros((n, n), dtype=np.int64)
    for i in range(n):
        for j in range(n):
            a[i, j] = SPT[i][j] if j in SPT[i] else 100
    return a.reshape(-1)


if __name__ == '__main__':
    store = {}
    for name, (ei, n) in cases().items():
        store[name + '/edge_index'] = ei
        store[name + '/n'] = np.array([n])
        store[name + '/attn_bias'] = reference_attn_bias(ei, n)
    np.savez_compressed(os.path.join(HERE, 'spd.npz'), **store)
    print('spd.npz', {k: v.shape for k, v in store.items() if k.endswith('attn_bias')})
