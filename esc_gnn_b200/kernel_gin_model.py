"""Graph-level `NestedGIN_eff` of the reference's kernel package (graph classification: CSL / EXP / TU style runs) --
constructor, forward contract and state_dict keys of /root/reference/kernel/gin.py:200-379, on the sm_100a kernels
(bag-embed, GINE aggregation, mean pooling, tcgen05 Linear, BatchNorm)."""
import torch
import torch.nn.functional as F

from . import ops
from .gine import GINEConv
from .graphcount_model import _mlp, _z_embedding
from .ops import Linear


class NestedGIN_eff(torch.nn.Module):
    def __init__(self, dataset, num_layers, hidden, use_z=False, use_rd=False, use_cycle=False, graph_pred=True,
                 use_id=None, dropout=0.2, multi_layer=False, edge_nest=False):
        super(NestedGIN_eff, self).__init__()
        if use_id is not None:
            raise NotImplementedError('use_id selects the non-efficient GINIDConvLayer path (kernel/gin.py:254-276)')
        self.use_rd = use_rd
        self.use_z = True
        self.graph_pred = graph_pred
        self.use_cycle = use_cycle
        self.use_id = use_id
        self.dropout = dropout
        self.multi_layer = multi_layer
        self.edge_nest = edge_nest
        self.z_initial = torch.nn.Embedding(1800, hidden)
        self.z_embedding = _z_embedding(hidden, dropout)
        input_dim = dataset.num_features
        self.conv1 = GINEConv(_mlp(input_dim, hidden, dropout), train_eps=True, edge_dim=hidden)
        self.convs = torch.nn.ModuleList()
        for _ in range(num_layers - 1):
            self.convs.append(GINEConv(_mlp(hidden, hidden, dropout), train_eps=True, edge_dim=hidden))
        self.lin1 = ops.Linear(num_layers * hidden, hidden)
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
            raise NotImplementedError('dense edge_pos is the legacy slow path (kernel/gin.py:335-338)')
        index = ops.graph_index(data)
        x, edge_index = data.x, data.edge_index
        z_emb = self.z_embedding(ops.bag_embed_data(self.z_initial.weight, data, index))
        x = self.conv1(x, edge_index, z_emb, index)
        xs = [x]
        for conv in self.convs:
            x = conv(x, edge_index, z_emb, index)
            xs += [x]
        x = torch.cat(xs, dim=1)
        if self.graph_pred:
            x = ops.global_mean_pool(x, index)
        x = self.lin1(x)
        if x.size()[0] > 1:
            x = self.bn_lin1(x)
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = F.relu(x)
        x = self.lin2(x)
        return x if self.use_cycle else F.log_softmax(x, dim=-1)
