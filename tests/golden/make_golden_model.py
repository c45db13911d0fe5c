"""Golden fixtures for the model step, produced by the reference's OWN classes.

Run in the build container only:  python tests/golden/make_golden_model.py
The `NestedGIN_eff` / `GNN` classes are AST-extracted from the unmodified files under /root/reference
(run_graphcount.py, zinc_models.py, ogb_mol_gnn.py; the scripts execute argparse / dataset loading at import time and
import k_gnn, so they cannot be imported whole -- SURVEY.md Appendix D) and executed under the PyG stand-in of
tests/_pyg_shim with deterministic weights (tests/model_util.det_state).  Stored per case: predictions and loss of a
training-mode forward, eval-mode predictions, digests of every parameter gradient, and the loss trajectory of three
Adam steps (pins BatchNorm running statistics, the optimiser and the backward pass).
"""
import ast
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'tests', '_pyg_shim'))
sys.path.insert(0, ROOT)

from torch_geometric.nn import GINEConv, MessagePassing, global_add_pool, global_mean_pool  # noqa: E402 (stand-ins)

from esc_gnn_b200 import synth  # noqa: E402
from tests import model_util as MU  # noqa: E402
from oracle import model_ref  # noqa: E402


class BondEncoder(torch.nn.Module):           # ogb.graphproppred.mol_encoder.BondEncoder (third-party) stand-in
    def __init__(self, emb_dim):
        super().__init__()
        self.bond_embedding_list = torch.nn.ModuleList()
        for d in model_ref.BOND_DIMS:
            emb = torch.nn.Embedding(d, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            self.bond_embedding_list.append(emb)

    def forward(self, edge_attr):
        out = 0
        for i in range(edge_attr.shape[1]):
            out += self.bond_embedding_list[i](edge_attr[:, i])
        return out


def extract(path, names):
    tree = ast.parse(open(path).read())
    ns = dict(torch=torch, F=F, Linear=torch.nn.Linear, Sequential=torch.nn.Sequential, ReLU=torch.nn.ReLU,
              ELU=torch.nn.ELU, BN=torch.nn.BatchNorm1d, Dropout=torch.nn.Dropout, GINEConv=GINEConv,
              MessagePassing=MessagePassing, global_add_pool=global_add_pool, global_mean_pool=global_mean_pool,
              global_max_pool=None, BondEncoder=BondEncoder, get_atom_feature_dims=lambda: list(model_ref.ATOM_DIMS),
              np=np)
    seen = set()
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names and node.name not in seen:
            seen.add(node.name)                # first definition of each name
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
    return ns


class _DS(object):
    num_classes = 1
    num_features = 10


def build_reference_model(variant, kw):
    if variant == 'count':
        ns = extract('/root/reference/run_graphcount.py', {'NestedGIN_eff'})
        return ns['NestedGIN_eff'](_DS(), kw['num_layers'], kw['hidden'], use_rd=True, graph_pred=False, dropout=0,
                                   edge_nest=True, use_cycle=True)          # run_graphcount.py:465
    if variant == 'zinc':
        ns = extract('/root/reference/zinc_models.py', {'NestedGIN_eff'})
        return ns['NestedGIN_eff'](_DS(), kw['num_layers'])                  # run_zinc.py:241-257
    if variant == 'kgin':
        ns = extract('/root/reference/kernel/gin.py', {'NestedGIN_eff'})

        class _TU(object):
            num_features, num_classes = synth.KGIN_FEATURES, synth.KGIN_CLASSES
        return ns['NestedGIN_eff'](_TU(), kw['num_layers'], kw['hidden'], use_rd=True, dropout=0)
    if variant == 'qm9':
        ns = extract('/root/reference/qm9_models.py', {'NestedGIN_eff'})

        class _QM9(object):
            num_features = synth.QM9_FEATURES
        return ns['NestedGIN_eff'](_QM9(), kw['num_layers'])                 # run_qm9.py:262-270
    ns = extract('/root/reference/ogb_mol_gnn.py', {'GNN', 'AtomEncoder', 'GINConv_eff', 'GNN_node_efficient',
                                                      'center_pool', 'center_pool_virtual'})
    return ns['GNN']('ogbg-molhiv', kw['num_tasks'], num_layer=kw['num_layer'], emb_dim=kw['emb_dim'],
                     gnn_type='gin_eff', virtual_node=kw['virtual_node'], residual=kw['residual'],
                     drop_ratio=kw['drop_ratio'])                            # run_ogb_mol.py:432-434


class _Batch(object):                           # what batch.py hands the model: attribute bag with .to()
    def __init__(self, b):
        self.__dict__.update(b.__dict__)

    def to(self, device):
        return self

    def __contains__(self, key):
        return key in self.__dict__


def main():
    torch.set_num_threads(4)
    path = os.path.join(HERE, 'model.npz')
    store = {}
    if os.path.exists(path) and '--all' not in sys.argv:       # default: keep the committed cases, add the missing ones
        with np.load(path) as old:
            store = {k: old[k] for k in old.files}
    for name, (variant, config, count, kw) in MU.MODEL_CASES.items():
        if name + '/loss' in store:
            continue
        torch.manual_seed(0)
        model = build_reference_model(variant, kw)
        sd = MU.det_state(model.state_dict(), seed=1234)
        model.load_state_dict(sd)
        batch = _Batch(MU.ref_batch(config, 100, count))
        model.train()
        pred = model(batch)
        loss = MU.loss_fn(variant, pred, batch.y)
        loss.backward()
        store[name + '/pred_train'] = pred.detach().numpy().astype(np.float32)
        store[name + '/loss'] = np.array([loss.item()], dtype=np.float64)
        keys, dig = [], []
        for k, p in model.named_parameters():
            if p.grad is not None:
                keys.append(k); dig.append(MU.grad_digest(p.grad))
        store[name + '/grad_keys'] = np.array(keys)
        store[name + '/grad_digest'] = np.stack(dig)
        # running stats after that one training forward
        sd1 = model.state_dict()
        rk = sorted(k for k in sd1 if k.endswith('running_mean') or k.endswith('running_var'))
        store[name + '/running_keys'] = np.array(rk)
        store[name + '/running_digest'] = np.stack([MU.grad_digest(sd1[k]) for k in rk])
        model.eval()
        with torch.no_grad():
            store[name + '/pred_eval'] = model(batch).numpy().astype(np.float32)
        # three Adam steps from the deterministic state
        model.load_state_dict(sd)
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        traj = []
        for _ in range(3):
            opt.zero_grad()
            l = MU.loss_fn(variant, model(batch), batch.y)
            l.backward()
            opt.step()
            traj.append(l.item())
        store[name + '/adam_losses'] = np.array(traj, dtype=np.float64)
        model.eval()
        with torch.no_grad():
            store[name + '/pred_after_adam'] = model(batch).numpy().astype(np.float32)
        if name in MU.FP64_CASES:
            # fp64 truth of the SAME reference class: per-tensor ||g32 - g64||_2 is the yardstick the product's gradients are
            # measured with (tests: ||g_product - g64|| <= 3 * ||g32_reference - g64|| + floor)
            model.load_state_dict(sd)
            model.train()
            g32 = {}
            model.zero_grad()
            MU.loss_fn(variant, model(batch), batch.y).backward()
            for k, p in model.named_parameters():
                if p.grad is not None:
                    g32[k] = p.grad.detach().double().clone()
            model64 = build_reference_model(variant, kw).double()
            model64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()})
            model64.train()
            b64 = _Batch(MU.to_double(batch))
            pred64 = model64(b64)
            loss64 = MU.loss_fn(variant, pred64, b64.y) if variant != 'ogb' else \
                torch.nn.BCEWithLogitsLoss()(pred64[b64.y.view(pred64.shape) == b64.y.view(pred64.shape)],
                                             b64.y.view(pred64.shape)[b64.y.view(pred64.shape) == b64.y.view(pred64.shape)])
            loss64.backward()
            dig64, err32 = [], []
            for k in keys:
                g64 = dict(model64.named_parameters())[k].grad.detach()
                dig64.append(MU.grad_digest(g64))
                err32.append((g32[k] - g64).norm().item())
            store[name + '/grad64_digest'] = np.stack(dig64)
            store[name + '/grad_err32'] = np.array(err32, dtype=np.float64)
            store[name + '/loss64'] = np.array([loss64.item()], dtype=np.float64)
            store[name + '/pred64'] = pred64.detach().numpy()
        nparam = sum(p.numel() for p in model.parameters())
        print('%-12s params %8d  N %5d  E %6d  nnz %7d  loss %.6f  adam %s' % (
            name, nparam, batch.x.shape[0], batch.edge_index.shape[1], batch.pos_enc.numel(), loss.item(),
            ['%.5f' % t for t in traj]))
    np.savez_compressed(path, **store)
    print('model.npz %.1f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
