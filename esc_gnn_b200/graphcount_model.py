"""`NestedGIN_eff` for the counting benchmarks -- constructor, forward contract and state_dict keys of
/root/reference/run_graphcount.py:39-194, running on the sm_100a kernels (bag-embed, GINE aggregation, pooling)."""
import torch
import torch.nn.functional as F
from torch.nn import Dropout, ReLU, Sequential

from .ops import BatchNorm1d as BN
from .ops import Linear

from . import ops
from .gine import GINEConv


def _mlp(cin, hidden, dropout, act=ReLU):
    return Sequential(Linear(cin, hidden), Dropout(dropout), BN(hidden), act(), Linear(hidden, hidden),
                      Dropout(dropout), BN(hidden), act())


def _z_embedding(hidden, dropout, act=ReLU):
    return Sequential(Dropout(dropout), BN(hidden), act(), Linear(hidden, hidden), Dropout(dropout), BN(hidden), act())


class NestedGIN_eff(torch.nn.Module):
    def __init__(self, dataset, num_layers, hidden, use_z=False, use_rd=False, use_cycle=False, graph_pred=True,
                 use_id=None, dropout=0.2, multi_layer=False, edge_nest=False):
        super(NestedGIN_eff, self).__init__()
        if use_id is not None:
            raise NotImplementedError('use_id selects the non-efficient node_id path (run_graphcount.py:162-173)')
        self.use_rd = use_rd
        self.use_z = True
        self.graph_pred = graph_pred
        self.use_cycle = use_cycle
        self.use_id = use_id
        self.dropout = dropout
        self.multi_layer = multi_layer
        self.edge_nest = edge_nest
        input_dim = 10
        self.z_initial = torch.nn.Embedding(1800, hidden)
        self.z_embedding = _z_embedding(hidden, dropout)
        self.x_embedding = _mlp(input_dim, hidden, dropout)
        self.conv1 = GINEConv(_mlp(input_dim, hidden, dropout), train_eps=True, edge_dim=hidden)
        self.convs = torch.nn.ModuleList()
        for _ in range(num_layers - 1):
            self.convs.append(GINEConv(_mlp(hidden, hidden, dropout), train_eps=True, edge_dim=hidden))
        self.lin1 = ops.Linear(num_layers * hidden + hidden, hidden)
        self.bn_lin1 = ops.BatchNorm1d(hidden, eps=1e-5, momentum=0.1)
        self.lin2 = Linear(hidden, 1) if use_cycle else Linear(hidden, dataset.num_classes)

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

    def forward(self, data):
        data.to(self.lin1.weight.device)
        if hasattr(data, 'edge_pos'):
            raise NotImplementedError('dense edge_pos is the legacy slow path (run_graphcount.py:142-146)')
        index = ops.graph_index(data)
        x, edge_index = data.x, data.edge_index
        z_emb = self.z_embedding(ops.bag_embed_data(self.z_initial.weight, data, index))
        x = self.conv1(x, edge_index, z_emb, index)
        xs = [self.x_embedding(data.x), x]
        for conv in self.convs:
            x = conv(x, edge_index, z_emb, index)
            xs += [x]
        x = torch.cat(xs, dim=1)
        if self.graph_pred:
            x = ops.global_mean_pool(x, index)
        x = self.lin1(x)
        if x.size()[0] > 1:
            x = self.bn_lin1(x)
        x = F.dropout(F.relu(x), p=self.dropout, training=self.training)
        x = self.lin2(x)
        return x if self.use_cycle else F.log_softmax(x, dim=-1)
