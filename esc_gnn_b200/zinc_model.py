"""`NestedGIN_eff` for ZINC -- constructor, forward contract and state_dict keys of
/root/reference/zinc_models.py:504-611 (ELU, atom / bond type embeddings, sum-pool readout), on the sm_100a kernels."""
import torch
import torch.nn.functional as F
from torch.nn import ELU

from . import ops
from .gine import GINEConv
from .graphcount_model import _mlp, _z_embedding
from .ops import Linear


class NestedGIN_eff(torch.nn.Module):
    def __init__(self, dataset, num_layers, concat=False, use_pos=False, use_max_dist=False, RNI=False, **kwargs):
        super(NestedGIN_eff, self).__init__()
        self.use_z = True
        hidden = kwargs.pop('hidden', 256)          # the reference hard-codes 256 (:508); tests may shrink it
        dropout = 0.0
        self.dropout = dropout
        self.z_initial = torch.nn.Embedding(1800, hidden)
        self.z_embedding = _z_embedding(hidden, dropout, ELU)
        input_dim, edge_attr_dim = 32, 32
        self.conv1 = GINEConv(_mlp(input_dim, hidden, dropout, ELU), train_eps=True, edge_dim=hidden + edge_attr_dim)
        self.convs = torch.nn.ModuleList()
        for _ in range(num_layers - 1):
            self.convs.append(GINEConv(_mlp(hidden, hidden, dropout, ELU), train_eps=True,
                                       edge_dim=hidden + edge_attr_dim))
        self.lin1 = ops.Linear(num_layers * hidden, hidden)
        self.bn_lin1 = ops.BatchNorm1d(hidden, eps=1e-5, momentum=0.1)
        self.lin2 = Linear(hidden, 1)
        self.node_type_embedding = torch.nn.Embedding(100, 32)
        self.edge_type_embedding = torch.nn.Embedding(100, 32)

    def reset_parameters(self):
        for layer in self.z_embedding.children():
            if hasattr(layer, 'reset_parameters'):
                layer.reset_parameters()
        self.conv1.reset_parameters()
        for conv in self.convs:
            conv.reset_parameters()
        self.lin1.reset_parameters()
        self.bn_lin1.reset_parameters()
        self.lin2.reset_parameters()
        self.node_type_embedding.reset_parameters()
        self.edge_type_embedding.reset_parameters()

    def forward(self, data):
        data.to(self.lin1.weight.device)
        if hasattr(data, 'edge_pos'):
            raise NotImplementedError('dense edge_pos is the legacy slow path (zinc_models.py:584-587)')
        index = ops.graph_index(data)
        x, edge_index = self.node_type_embedding(data.x), data.edge_index
        z_emb = self.z_embedding(ops.bag_embed_data(self.z_initial.weight, data, index))
        z_emb = torch.cat((z_emb, self.edge_type_embedding(data.edge_attr)), dim=-1)
        x = self.conv1(x, edge_index, z_emb, index)
        xs = [x]
        for conv in self.convs:
            x = conv(x, edge_index, z_emb, index)
            xs += [x]
        x = ops.global_add_pool(torch.cat(xs, dim=1), index)
        x = self.lin1(x)
        if x.size()[0] > 1:
            x = self.bn_lin1(x)
        x = F.elu(F.dropout(x, p=self.dropout, training=self.training))
        return self.lin2(x)
