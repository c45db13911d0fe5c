"""Golden fixtures for the GraphGPS twin, from the UNMODIFIED reference sources, build container only:
    python tests/golden/make_golden_gps.py

* collation: `/root/reference/GraphGPS/graphgps/loader/batch.py` imported as-is under the PyG stand-in, fed graphs that carry an
  `attn_bias` key (what `create_subgraphs` of loader/utils_escgnn.py returns);
* edge-feature injection: gps_layer.py cannot be imported (graphgym / performer dependencies), so the statements that ARE the
  injection are lifted by `ast` out of `GPSLayer.__init__` (the `z_in`, `hidden`, `self.z_initial`, `self.z_embedding` assignments,
  gps_layer.py:169-183) and out of `GPSLayer.forward` (its first `if hasattr(batch, 'pos_index')` statement, :186-188) into a
  two-method module and executed with deterministic weights.
Stored in gps.npz: gps/batch/<key> and gps/inject/{edge_attr_in, edge_attr_out, grad_digest...}.
"""
import ast
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'tests', '_pyg_shim'))
sys.path.insert(0, ROOT)

from torch_geometric.data import Data  # noqa: E402 (stand-in)
from torch_geometric.nn import global_add_pool  # noqa: E402

from tests import model_util as MU  # noqa: E402
from tests.gps_cases import GPS_BATCH, GPS_INJECT, gps_graphs, inject_batch  # noqa: E402

SRC = '/root/reference/GraphGPS/graphgps/layer/gps_layer.py'


def reference_injection_class():
    tree = ast.parse(open(SRC).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'GPSLayer'][0]
    init = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == '__init__'][0]
    fwd = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == 'forward'][0]

    def targets(stmt):
        return [ast.unparse(t) for t in stmt.targets] if isinstance(stmt, ast.Assign) else []
    keep = [s for s in init.body if any(t in ('z_in', 'hidden', 'self.z_initial', 'self.z_embedding') for t in targets(s))]
    assert len(keep) == 4, [ast.unparse(s)[:40] for s in keep]
    first_if = [s for s in fwd.body if isinstance(s, ast.If)][0]
    assert 'pos_index' in ast.unparse(first_if.test)
    src = 'class Inject(nn.Module):\n    def __init__(self, dim_h, dropout):\n        super().__init__()\n'
    src += ''.join('        ' + ast.unparse(s).replace('\n', '\n        ') + '\n' for s in keep)
    src += '    def forward(self, batch):\n' + ''.join('        ' + line + '\n' for line in ast.unparse(first_if).split('\n'))
    src += '        return batch\n'
    ns = dict(nn=torch.nn, torch=torch, global_add_pool=global_add_pool)
    exec(src, ns)
    return ns['Inject'], src


def main():
    store = {}
    spec = importlib.util.spec_from_file_location('gps_loader_batch', '/root/reference/GraphGPS/graphgps/loader/batch.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    b = mod.Batch.from_data_list(gps_graphs(Data, *GPS_BATCH))
    for k in b.keys:
        v = b[k]
        if torch.is_tensor(v):
            store['gps/batch/%s' % k] = v.numpy()
    store['gps/batch/has_attn_bias_key'] = np.array([int('attn_bias' in b._store if hasattr(b, '_store') else 0)])
    store['gps/batch/attn_bias_is_none'] = np.array([int(b.attn_bias is None if hasattr(b, 'attn_bias') else 1)])
    Inject, src = reference_injection_class()
    dim_h, config, start, count = GPS_INJECT
    m = Inject(dim_h, 0.0)
    m.load_state_dict(MU.det_state(m.state_dict(), seed=4321))
    m.train()
    batch = inject_batch(config, start, count, dim_h)
    ea_in = batch.edge_attr.clone().requires_grad_(True)
    batch.edge_attr = ea_in * 1.0
    out = m(batch).edge_attr
    (out * torch.linspace(-1, 1, out.numel()).view_as(out)).sum().backward()
    store['gps/inject/edge_attr_out'] = out.detach().numpy()
    keys, dig = [], []
    for k, p in m.named_parameters():
        keys.append(k); dig.append(MU.grad_digest(p.grad))
    store['gps/inject/grad_keys'] = np.array(keys)
    store['gps/inject/grad_digest'] = np.stack(dig)
    store['gps/inject/source'] = np.array([src])
    path = os.path.join(HERE, 'gps.npz')
    np.savez_compressed(path, **store)
    print(src)
    print('gps.npz: %d arrays, %.1f KB' % (len(store), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
