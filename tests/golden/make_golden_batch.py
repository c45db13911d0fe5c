"""Golden fixtures for collation (B1) and the `Distance` edge transform (D1), produced by the UNMODIFIED reference files.

Run in the build container only (needs the read-only reference mount):
    python tests/golden/make_golden_batch.py
Imports `/root/reference/batch.py` and `/root/reference/distance.py` as they are, under the PyG stand-in of
`tests/_pyg_shim` (torch_geometric is not installable here), and freezes their outputs into `batch.npz`:

    batch/<case>/<key>        every tensor key of `Batch.from_data_list(list)` (+ `batch`, `<key>_batch` of follow_batch)
    dist/<case>/in_*          inputs of one `Distance(**kw)(data)` call (pos, edge_index, edge_attr, original_* keys)
    dist/<case>/edge_attr     its output (and `original_edge_attr` when the original_* keys are present)

The collation inputs are NOT stored: they are `tests.model_util.graph_dicts(config, start, count)` (synthetic graphs +
C-oracle encodings, deterministic), rebuilt by the test; `batch/<case>/meta` holds (config, start, count, variantflag).
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
import batch as REF_BATCH  # noqa: E402     (the unmodified reference)
import distance as REF_DIST  # noqa: E402   (the unmodified reference)

from tests.batch_cases import BATCH_CASES, DIST_CASES, batch_inputs, dist_inputs, follow  # noqa: E402


def main():
    store = {}
    for name, config, start, count, variant in BATCH_CASES:
        b = REF_BATCH.Batch.from_data_list(batch_inputs(Data, config, start, count, variant), follow_batch=follow(variant))
        store['batch/%s/meta' % name] = np.array([config, start, count, variant], dtype=np.int64)
        for k in b.keys:
            v = b[k]
            if torch.is_tensor(v):
                store['batch/%s/%s' % (name, k)] = v.numpy()
        store['batch/%s/num_graphs' % name] = np.array([b.num_graphs], dtype=np.int64)
    for i, (name, kw, has_attr, one_d, original) in enumerate(DIST_CASES):
        inp = dist_inputs(100 + i, has_attr, one_d, original)
        d = Data(x=torch.ones(inp['pos'].size(0), 1), **{k: v.clone() for k, v in inp.items()})
        o = REF_DIST.Distance(**kw)(d)
        for k, v in inp.items():
            store['dist/%s/in_%s' % (name, k)] = v.numpy()
        store['dist/%s/edge_attr' % name] = o.edge_attr.numpy()
        if original:
            store['dist/%s/original_edge_attr' % name] = o.original_edge_attr.numpy()
    path = os.path.join(HERE, 'batch.npz')
    np.savez_compressed(path, **store)
    print('batch.npz: %d arrays, %.1f KB' % (len(store), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
