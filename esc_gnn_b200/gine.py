"""GINE-style convolutions on the sm_100a aggregation kernel.

`GINEConv` keeps PyG 2.0.4's constructor and parameter names (`nn.*`, `eps`, `lin.*`) as the reference uses it
(/root/reference/run_graphcount.py:77-109, zinc_models.py:527-557; semantics restated in-tree at
GraphGPS/graphgps/layer/gine_conv_layer.py:18-35,56-84).  `GINConv_eff` mirrors /root/reference/ogb_mol_gnn.py:323-358.
Message, ReLU, segmented sum and the (1+eps) residual are one kernel (ops.gine_aggregate): no [E,C] message tensor,
no atomics.
"""
import torch

from . import ops


def _reset(module):
    if hasattr(module, 'reset_parameters'):
        module.reset_parameters()
    else:
        for m in module.children():
            _reset(m)


class GINEConv(torch.nn.Module):
    def __init__(self, nn, eps=0., train_eps=False, edge_dim=None, **kwargs):
        super(GINEConv, self).__init__()
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer('eps', torch.Tensor([eps]))
        if edge_dim is not None:
            first = self.nn[0]
            in_channels = first.in_features if hasattr(first, 'in_features') else first.in_channels
            self.lin = ops.Linear(edge_dim, in_channels)
        else:
            self.lin = None
        self.reset_parameters()

    def reset_parameters(self):
        _reset(self.nn)
        self.eps.data.fill_(self.initial_eps)
        if self.lin is not None:
            self.lin.reset_parameters()

    def forward(self, x, edge_index, edge_attr=None, index=None):
        if index is None:
            index = ops.GraphIndex(edge_index, x.size(0))
        if self.lin is None and x.size(-1) != edge_attr.size(-1):
            raise ValueError("Node and edge feature dimensionalities do not match. Consider setting the 'edge_dim' "
                             "attribute of 'GINEConv'")
        e = self.lin(edge_attr) if self.lin is not None else edge_attr
        return self.nn(ops.gine_aggregate(x, e, self.eps, index))

    def __repr__(self):
        return '{}(nn={})'.format(self.__class__.__name__, self.nn)


class SumEmbedding(torch.nn.Module):
    """Sum of per-column embeddings: ogb's AtomEncoder / BondEncoder (ogb_mol_gnn.py:264-282; ogb is third-party)."""
    def __init__(self, dims, emb_dim, list_name):
        super(SumEmbedding, self).__init__()
        lst = torch.nn.ModuleList()
        for d in dims:
            emb = torch.nn.Embedding(d, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            lst.append(emb)
        setattr(self, list_name, lst)
        self._list_name = list_name

    def forward(self, x):
        out = 0
        for i, emb in enumerate(getattr(self, self._list_name)):
            out = out + emb(x[:, i])
        return out


class GINConv_eff(torch.nn.Module):
    def __init__(self, dataset, emb_dim, bond_dims=(5, 6, 2)):
        super(GINConv_eff, self).__init__()
        self.mlp = torch.nn.Sequential(ops.Linear(emb_dim, 2 * emb_dim), ops.BatchNorm1d(2 * emb_dim),
                                       torch.nn.ReLU(), ops.Linear(2 * emb_dim, emb_dim))
        self.eps = torch.nn.Parameter(torch.Tensor([0]))
        if dataset.startswith('ogbg-mol'):
            self.edge_encoder = SumEmbedding(bond_dims, emb_dim, 'bond_embedding_list')
        elif dataset.startswith('ogbg-ppa'):
            self.edge_encoder = ops.Linear(7, emb_dim)
        self.edge_encoder_pos = ops.Linear(emb_dim, emb_dim)

    def forward(self, x, edge_index, edge_attr, edge_pos, index=None):
        if index is None:
            index = ops.GraphIndex(edge_index, x.size(0))
        e = self.edge_encoder(edge_attr) + self.edge_encoder_pos(edge_pos)
        return self.mlp(ops.gine_aggregate(x, e, self.eps, index))
