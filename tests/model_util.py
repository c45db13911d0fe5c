"""Shared by tests/golden/make_golden_model.py (reference side) and the model parity tests: deterministic weights,
deterministic synthetic batches (synth graphs + C-oracle encodings + reference collation rules)."""
import math

import numpy as np
import torch

from esc_gnn_b200 import synth
from oracle import c_oracle, model_ref


def det_state(state_dict, seed):
    """Deterministic replacement values for every entry of a state_dict (same on every machine with this torch)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(state_dict.keys()):
        v = state_dict[k]
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros_like(v)
        elif k.endswith('running_var'):
            out[k] = 1.0 + 0.2 * torch.rand(v.shape, generator=g)
        elif k.endswith('running_mean'):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('eps'):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif v.dim() >= 2:
            out[k] = torch.randn(v.shape, generator=g) / math.sqrt(v.shape[1])
        elif k.endswith('weight'):        # BatchNorm scale
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
    return out


def graph_dicts(config, start, count, h=None, use_rd=None, self_loop=None):
    """Per-graph dicts (torch tensors) as the reference dataset would hold them after the pre_transform."""
    fl = dict(synth.ENCODER_FLAGS[config])
    if h is not None:
        fl['h'] = h
    if use_rd is not None:
        fl['use_rd'] = use_rd
    if self_loop is not None:
        fl['self_loop'] = self_loop
    out = []
    for i in range(start, start + count):
        g = synth.make_graph(config, i)
        eo, pe, pi, pb = c_oracle.encode_graph(g['edge_index'], g['num_nodes'], fl['h'], fl['use_rd'], fl['self_loop'])
        d = dict(x=torch.as_tensor(g['x']), edge_index=torch.as_tensor(eo), y=torch.as_tensor(g['y']),
                 pos_enc=torch.as_tensor(pe), pos_index=torch.as_tensor(pi), pos_batch=torch.as_tensor(pb),
                 num_nodes=g['num_nodes'])
        if 'edge_attr' in g:
            ea = torch.as_tensor(g['edge_attr'])
            if fl['self_loop']:      # E1 on attributes: (no loops in synth graphs) append N rows of ones
                ea = torch.cat([ea, ea.new_full((g['num_nodes'], ) + tuple(ea.shape[1:]), 1)], 0)
            d['edge_attr'] = ea
        if 'pos' in g:       # QM9 flow: Distance(norm=True) runs as a load-time transform AFTER the pre_transform
            d['pos'], d['node_type'] = torch.as_tensor(g['pos']), torch.as_tensor(g['node_type'])
            d['edge_attr'] = model_ref.distance_transform(d['edge_index'], d['pos'], d['edge_attr'])
        out.append(d)
    return out


def ref_batch(config, start, count, **kw):
    return model_ref.collate(graph_dicts(config, start, count, **kw))


def grad_digest(t):
    t = t.detach().double().flatten().cpu()
    head = t[:6].tolist() + [0.0] * max(0, 6 - t.numel())
    return np.array([t.sum().item(), t.abs().sum().item(), t.abs().max().item(), t.norm().item()] + head)


MODEL_CASES = {
    # name: (variant, config, graphs, ctor kwargs)
    'count_h64': ('count', 1, 24, dict(num_layers=3, hidden=64)),
    'count_h256': ('count', 1, 16, dict(num_layers=5, hidden=256)),
    'zinc': ('zinc', 2, 48, dict(num_layers=5)),
    'zinc_l2': ('zinc', 2, 3, dict(num_layers=2)),
    'ogb': ('ogb', 4, 12, dict(num_tasks=1, num_layer=3, emb_dim=64, drop_ratio=0.0, virtual_node=True, residual=True)),
    'qm9': ('qm9', 6, 20, dict(num_layers=3)),
    'kgin': ('kgin', 7, 20, dict(num_layers=3, hidden=64)),            # kernel/gin.py graph-classification variant
    'ogb_full': ('ogb', 4, 8, dict(num_tasks=1, num_layer=6, emb_dim=300, drop_ratio=0.0, virtual_node=True, residual=False)),
    # BASELINE.json shapes (configs[0..3]): the reference's own batch sizes, depths and widths; these cases also carry fp64
    # gradient statistics of the reference class (`/grad64_digest`, `/grad_err32`)
    'count_cfg1': ('count', 1, 128, dict(num_layers=5, hidden=256)),       # run_graphcount.py:465, batch 128, h=3
    'zinc_cfg2': ('zinc', 2, 256, dict(num_layers=5)),                     # run_zinc.py:56 batch 256, h=3
    'count_cfg3': ('count', 3, 32, dict(num_layers=5, hidden=256)),        # count_graphlet, h=4, batch 32
    'ogb_cfg4': ('ogb', 4, 32, dict(num_tasks=1, num_layer=6, emb_dim=300, drop_ratio=0.0, virtual_node=True, residual=False)),
}
FP64_CASES = ('count_cfg1', 'zinc_cfg2', 'count_cfg3', 'ogb_cfg4')


def to_double(batch):
    """The same batch with every floating tensor in float64 (fp64 gradient truth runs)."""
    import copy
    b = copy.copy(batch)
    for k, v in list(b.__dict__.items()):
        if torch.is_tensor(v) and v.is_floating_point():
            setattr(b, k, v.double())
    return b


def loss_fn(variant, pred, y):
    if variant == 'ogb':
        y = y.to(torch.float32).view(pred.shape)
        lab = y == y
        return torch.nn.BCEWithLogitsLoss()(pred.to(torch.float32)[lab], y[lab])      # run_ogb_mol.py:58-74
    if variant == 'qm9':
        return torch.nn.functional.mse_loss(pred, y.view(-1))                          # run_qm9.py:348
    if variant == 'kgin':
        return torch.nn.functional.nll_loss(pred, y.view(-1))                          # kernel/train_eval.py (F.nll_loss on log_softmax)
    return torch.nn.L1Loss()(pred, y.view(-1, 1))                                      # run_graphcount.py:494-500


def build_oracle_model(variant, kw):
    if variant == 'count':
        return model_ref.NestedGINEffCount(kw['num_layers'], kw['hidden'])
    if variant == 'zinc':
        return model_ref.NestedGINEffZinc(kw['num_layers'])
    if variant == 'qm9':
        return model_ref.NestedGINEffQM9(kw['num_layers'], synth.QM9_FEATURES)
    if variant == 'kgin':
        return model_ref.NestedGINEffKernel(kw['num_layers'], kw['hidden'], synth.KGIN_FEATURES, synth.KGIN_CLASSES)
    return model_ref.GNNOgbEff(kw['num_tasks'], kw['num_layer'], kw['emb_dim'], kw['virtual_node'], kw['residual'],
                               kw['drop_ratio'])
