"""OGB molecule models of the efficient path: `GNN(gnn_type='gin_eff')` -> `GNN_node_efficient` + `GINConv_eff`,
with the constructors, forward contract and state_dict keys of /root/reference/ogb_mol_gnn.py:66-117,252-261
(GNN), :264-282 (AtomEncoder), :614-792 (GNN_node_efficient), on the sm_100a kernels."""
import torch
import torch.nn.functional as F

from . import ops
from .gine import GINConv_eff, SumEmbedding
from .graphcount_model import _z_embedding

ATOM_DIMS = (119, 4, 12, 12, 10, 6, 6, 2, 2)     # ogb get_atom_feature_dims() (third-party)


class AtomEncoder(SumEmbedding):
    def __init__(self, emb_dim, dims=ATOM_DIMS):
        super(AtomEncoder, self).__init__(dims, emb_dim, 'atom_embedding_list')


class GNN_node_efficient(torch.nn.Module):
    def __init__(self, dataset, num_layer, emb_dim, drop_ratio=0.5, JK="last", residual=False, gnn_type='gin',
                 virtual_node=True, use_rd=False, adj_dropout=0, skip_node_encoder=False, use_rp=None,
                 center_pool_virtual=False, RNI=False):
        super(GNN_node_efficient, self).__init__()
        if gnn_type not in ('gin', 'gin_eff') or center_pool_virtual or RNI or skip_node_encoder:
            raise NotImplementedError('only the gin_eff configuration of the efficient path is built')
        self.num_layer, self.drop_ratio, self.JK = num_layer, drop_ratio, JK
        self.residual, self.virtual_node = residual, virtual_node
        self.z_initial = torch.nn.Embedding(1800, emb_dim)
        self.z_embedding = _z_embedding(emb_dim, drop_ratio)
        self.node_encoder = AtomEncoder(emb_dim)
        if self.virtual_node:
            self.virtualnode_embedding = torch.nn.Embedding(1, emb_dim)
            torch.nn.init.constant_(self.virtualnode_embedding.weight.data, 0)
        self.convs = torch.nn.ModuleList()
        self.batch_norms = torch.nn.ModuleList()
        for _ in range(num_layer):
            self.convs.append(GINConv_eff(dataset, emb_dim))
            self.batch_norms.append(ops.BatchNorm1d(emb_dim))
        if self.virtual_node:
            self.mlp_virtualnode_list = torch.nn.ModuleList()
            for _ in range(num_layer - 1):
                self.mlp_virtualnode_list.append(torch.nn.Sequential(
                    ops.Linear(emb_dim, 2 * emb_dim), ops.BatchNorm1d(2 * emb_dim), torch.nn.ReLU(),
                    ops.Linear(2 * emb_dim, emb_dim), ops.BatchNorm1d(emb_dim), torch.nn.ReLU()))

    def forward(self, batched_data, index=None):
        if hasattr(batched_data, 'edge_pos'):
            raise NotImplementedError('dense edge_pos is the legacy slow path (ogb_mol_gnn.py:710-713)')
        if index is None:
            index = ops.graph_index(batched_data)
        x, edge_index, edge_attr = batched_data.x, batched_data.edge_index, batched_data.edge_attr
        if self.virtual_node:
            vn = self.virtualnode_embedding.weight.expand(index.num_graphs, -1)
        h_list = [self.node_encoder(x)]
        z_emb = self.z_embedding(ops.bag_embed_data(self.z_initial.weight, batched_data, index))
        batch = batched_data.batch
        for layer in range(self.num_layer):
            if self.virtual_node:
                h_list[layer] = h_list[layer] + vn[batch]
            h = self.convs[layer](h_list[layer], edge_index, edge_attr, z_emb, index)
            h = self.batch_norms[layer](h)
            if layer == self.num_layer - 1:
                h = F.dropout(h, self.drop_ratio, training=self.training)
            else:
                h = F.dropout(F.relu(h), self.drop_ratio, training=self.training)
            if self.residual:
                h = h + h_list[layer]
            h_list.append(h)
            if self.virtual_node and layer < self.num_layer - 1:
                tmp = ops.global_add_pool(h_list[layer], index) + vn
                upd = F.dropout(self.mlp_virtualnode_list[layer](tmp), self.drop_ratio, training=self.training)
                vn = vn + upd if self.residual else upd
        if self.JK == "last":
            return h_list[-1]
        out = 0
        for layer in range(self.num_layer):
            out = out + h_list[layer]
        return out


class GNN(torch.nn.Module):
    def __init__(self, dataset, num_tasks, num_layer=5, emb_dim=300, gnn_type='gin', virtual_node=True, residual=False,
                 drop_ratio=0.5, JK="last", graph_pooling="mean", subgraph_pooling="mean", use_rd=False, use_rp=None,
                 RNI=False, deg_sub=None, deg_graph=None, **kwargs):
        super(GNN, self).__init__()
        if gnn_type != 'gin_eff':
            raise NotImplementedError("only gnn_type='gin_eff' (the efficient path) is built")
        if graph_pooling not in ('mean', 'sum'):
            raise NotImplementedError('graph_pooling must be mean or sum')
        self.num_layer, self.drop_ratio, self.JK, self.emb_dim = num_layer, drop_ratio, JK, emb_dim
        self.num_tasks, self.graph_pooling = num_tasks, graph_pooling
        self.gnn_node = GNN_node_efficient(dataset, num_layer, emb_dim, JK=JK, drop_ratio=drop_ratio,
                                           residual=residual, gnn_type=gnn_type, virtual_node=virtual_node,
                                           use_rd=use_rd, use_rp=use_rp, RNI=RNI)
        self.graph_pred_linear = ops.Linear(emb_dim, num_tasks)

    def forward(self, data, x=None, edge_index=None, edge_attr=None, batch=None, perturb=None):
        if perturb is not None:
            raise NotImplementedError('perturb (FLAG) is outside the hot path')
        data.to(self.graph_pred_linear.weight.device)
        index = ops.graph_index(data)
        h = self.gnn_node(data, index)
        pooled = ops.global_mean_pool(h, index) if self.graph_pooling == 'mean' else ops.global_add_pool(h, index)
        return self.graph_pred_linear(pooled)
