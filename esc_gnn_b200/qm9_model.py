"""`NestedGIN_eff` for QM9 -- constructor, forward contract and state_dict keys of /root/reference/qm9_models.py:25-139
(ReLU, continuous node features + 3-D positions with an additive node-type embedding, bond one-hot + `Distance` column
as continuous edge attributes, mean-pool readout), on the sm_100a kernels.  SURVEY.md section 8(f) N3."""
import torch
import torch.nn.functional as F
from torch.nn import ReLU

from . import ops
from .gine import GINEConv
from .graphcount_model import _mlp, _z_embedding


class NestedGIN_eff(torch.nn.Module):
    def __init__(self, dataset, num_layers, concat=False, use_pos=False, edge_attr_dim=5, use_max_dist=False, RNI=False,
                 **kwargs):
        super(NestedGIN_eff, self).__init__()
        self.use_z = True
        hidden = kwargs.pop('hidden', 256)          # the reference hard-codes 256 (:29); tests may shrink it
        dropout = 0.0
        self.dropout = dropout
        self.z_initial = torch.nn.Embedding(1800, hidden)
        self.z_embedding = _z_embedding(hidden, dropout, ReLU)
        input_dim = dataset.num_features + 3
        self.conv1 = GINEConv(_mlp(input_dim, hidden, dropout, ReLU), train_eps=True, edge_dim=hidden + edge_attr_dim)
        self.convs = torch.nn.ModuleList()
        for _ in range(num_layers - 1):
            self.convs.append(GINEConv(_mlp(hidden, hidden, dropout, ReLU), train_eps=True,
                                       edge_dim=hidden + edge_attr_dim))
        self.lin1 = ops.Linear(num_layers * hidden, hidden)
        self.bn_lin1 = ops.BatchNorm1d(hidden, eps=1e-5, momentum=0.1)
        self.lin2 = ops.Linear(hidden, 1)
        self.node_type_embedding = torch.nn.Embedding(5, input_dim)

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

    def forward(self, data):
        data.to(self.lin1.weight.device)
        if hasattr(data, 'edge_pos'):
            raise NotImplementedError('dense edge_pos is the legacy slow path (qm9_models.py:109-112)')
        index = ops.graph_index(data)
        x, edge_index = torch.cat([data.x, data.pos], 1), data.edge_index
        x = x + self.node_type_embedding(data.node_type)
        z_emb = self.z_embedding(ops.bag_embed_data(self.z_initial.weight, data, index))
        z_emb = torch.cat((z_emb, data.edge_attr), dim=-1)
        x = self.conv1(x, edge_index, z_emb, index)
        xs = [x]
        for conv in self.convs:
            x = conv(x, edge_index, z_emb, index)
            xs += [x]
        x = ops.global_mean_pool(torch.cat(xs, dim=1), index)
        x = self.lin1(x)
        if x.size()[0] > 1:
            x = self.bn_lin1(x)
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = F.relu(x)
        x = self.lin2(x)
        return x.view(-1)
