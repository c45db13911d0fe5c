"""ORACLE (test infrastructure, NOT product code) -- plain-PyTorch fp32 restatement of the NestedGIN_eff models.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module.  Same `state_dict` keys and shapes as the reference classes so weights move freely between the reference,
this oracle and the product modules.

Follows (paths under /root/reference):
  * count  : run_graphcount.py:39-194          (NestedGIN_eff, node-level, ReLU, x_embedding JK)
  * ZINC   : zinc_models.py:504-611            (NestedGIN_eff, ELU, type embeddings, global_add_pool readout)
  * QM9    : qm9_models.py:25-139              (NestedGIN_eff, continuous edge attributes, global_mean_pool readout)
             with distance.py:25-65 as the load-time edge transform
  * OGB    : ogb_mol_gnn.py:66-117,252-261 (GNN), :264-282 (AtomEncoder), :323-358 (GINConv_eff),
             :614-792 (GNN_node_efficient: virtual node, residual, BN + dropout)
  * GINEConv semantics (PyG 2.0.4; third-party, not under /root/reference): restated from the in-tree twin
    GraphGPS/graphgps/layer/gine_conv_layer.py:18-35,49-84
  * pooling: PyG global_add_pool / global_mean_pool (size = batch.max()+1; mean divides by count clamped >= 1)
  * collation: batch.py:25-149 (pos_batch += per-graph max+1, pos_enc / pos_index not offset)

Parity status: PINNED against outputs of the reference classes themselves (AST-extracted from the unmodified
files and executed under the PyG stand-in by tests/golden/make_golden_model.py; fixtures tests/golden/model_*.npz).
"""
import torch
import torch.nn.functional as F
from torch.nn import BatchNorm1d as BN
from torch.nn import Dropout, ELU, Linear, ReLU, Sequential

ATOM_DIMS = (119, 4, 12, 12, 10, 6, 6, 2, 2)     # ogb.utils.features.get_atom_feature_dims() (third-party, ogb 1.3.3)
BOND_DIMS = (5, 6, 2)                            # ogb get_bond_feature_dims()


def global_add_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    return x.new_zeros((size, ) + tuple(x.shape[1:])).index_add_(0, batch, x)


def global_mean_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    s = global_add_pool(x, batch, size)
    cnt = torch.zeros(size, dtype=x.dtype, device=x.device).index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return s / cnt.clamp(min=1).view(-1, 1)


def bag_embed(weight, pos_index, pos_enc, pos_batch):
    """z0[e] = sum_k pos_enc[k] * W[pos_index[k]] over pos_batch[k] == e   (run_graphcount.py:155)."""
    return global_add_pool(weight[pos_index] * pos_enc.view(-1, 1), pos_batch)


class GINEConv(torch.nn.Module):
    def __init__(self, nn, eps=0., train_eps=False, edge_dim=None):
        super().__init__()
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer('eps', torch.Tensor([eps]))
        self.lin = Linear(edge_dim, nn[0].in_features) if edge_dim is not None else None

    def forward(self, x, edge_index, edge_attr):
        e = self.lin(edge_attr) if self.lin is not None else edge_attr
        msg = (x[edge_index[0]] + e).relu()
        out = torch.zeros_like(x).index_add_(0, edge_index[1], msg)
        return self.nn(out + (1 + self.eps) * x)


def _mlp(cin, hidden, dropout, act):
    return Sequential(Linear(cin, hidden), Dropout(dropout), BN(hidden), act(), Linear(hidden, hidden),
                      Dropout(dropout), BN(hidden), act())


def _z_embedding(hidden, dropout, act):
    return Sequential(Dropout(dropout), BN(hidden), act(), Linear(hidden, hidden), Dropout(dropout), BN(hidden), act())


class NestedGINEffCount(torch.nn.Module):
    """run_graphcount.py:39-194 as constructed at :465 (use_id None, use_cycle=True -> lin2: hidden -> 1)."""
    def __init__(self, num_layers, hidden, graph_pred=False, dropout=0.0, use_cycle=True, num_classes=1):
        super().__init__()
        self.graph_pred, self.dropout, self.use_cycle = graph_pred, dropout, use_cycle
        self.z_initial = torch.nn.Embedding(1800, hidden)
        self.z_embedding = _z_embedding(hidden, dropout, ReLU)
        self.x_embedding = _mlp(10, hidden, dropout, ReLU)
        self.conv1 = GINEConv(_mlp(10, hidden, dropout, ReLU), train_eps=True, edge_dim=hidden)
        self.convs = torch.nn.ModuleList(
            [GINEConv(_mlp(hidden, hidden, dropout, ReLU), train_eps=True, edge_dim=hidden)
             for _ in range(num_layers - 1)])
        self.lin1 = Linear(num_layers * hidden + hidden, hidden)
        self.bn_lin1 = BN(hidden, eps=1e-5, momentum=0.1)
        self.lin2 = Linear(hidden, 1 if use_cycle else num_classes)

    def forward(self, data):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        z = self.z_embedding(bag_embed(self.z_initial.weight, data.pos_index, data.pos_enc, data.pos_batch))
        x = self.conv1(x, edge_index, z)
        xs = [self.x_embedding(data.x), x]
        for conv in self.convs:
            x = conv(x, edge_index, z)
            xs += [x]
        x = global_mean_pool(torch.cat(xs, dim=1), batch) if self.graph_pred else torch.cat(xs, dim=1)
        x = self.lin1(x)
        if x.size(0) > 1:
            x = self.bn_lin1(x)
        x = F.dropout(F.relu(x), p=self.dropout, training=self.training)
        x = self.lin2(x)
        return x if self.use_cycle else F.log_softmax(x, dim=-1)


class NestedGINEffKernel(torch.nn.Module):
    """kernel/gin.py:200-379 (graph classification): use_id None; JK = cat of the layer outputs, mean pooling, log_softmax."""
    def __init__(self, num_layers, hidden, num_features, num_classes, dropout=0.0, graph_pred=True, use_cycle=False):
        super().__init__()
        self.graph_pred, self.dropout, self.use_cycle = graph_pred, dropout, use_cycle
        self.z_initial = torch.nn.Embedding(1800, hidden)
        self.z_embedding = _z_embedding(hidden, dropout, ReLU)
        self.conv1 = GINEConv(_mlp(num_features, hidden, dropout, ReLU), train_eps=True, edge_dim=hidden)
        self.convs = torch.nn.ModuleList(
            [GINEConv(_mlp(hidden, hidden, dropout, ReLU), train_eps=True, edge_dim=hidden) for _ in range(num_layers - 1)])
        self.lin1 = Linear(num_layers * hidden, hidden)
        self.bn_lin1 = BN(hidden, eps=1e-5, momentum=0.1)
        self.lin2 = Linear(hidden, 1 if use_cycle else num_classes)

    def forward(self, data):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        z = self.z_embedding(bag_embed(self.z_initial.weight, data.pos_index, data.pos_enc, data.pos_batch))
        x = self.conv1(x, edge_index, z)
        xs = [x]
        for conv in self.convs:
            x = conv(x, edge_index, z)
            xs += [x]
        x = global_mean_pool(torch.cat(xs, dim=1), batch) if self.graph_pred else torch.cat(xs, dim=1)
        x = self.lin1(x)
        if x.size(0) > 1:
            x = self.bn_lin1(x)
        x = F.relu(F.dropout(x, p=self.dropout, training=self.training))
        x = self.lin2(x)
        return x if self.use_cycle else F.log_softmax(x, dim=-1)


class NestedGINEffZinc(torch.nn.Module):
    """zinc_models.py:504-611 (hidden 256 and dropout 0 are hard-coded there; parameterised here for small tests)."""
    def __init__(self, num_layers, hidden=256, dropout=0.0):
        super().__init__()
        self.dropout = dropout
        self.z_initial = torch.nn.Embedding(1800, hidden)
        self.z_embedding = _z_embedding(hidden, dropout, ELU)
        self.conv1 = GINEConv(_mlp(32, hidden, dropout, ELU), train_eps=True, edge_dim=hidden + 32)
        self.convs = torch.nn.ModuleList(
            [GINEConv(_mlp(hidden, hidden, dropout, ELU), train_eps=True, edge_dim=hidden + 32)
             for _ in range(num_layers - 1)])
        self.lin1 = Linear(num_layers * hidden, hidden)
        self.bn_lin1 = BN(hidden, eps=1e-5, momentum=0.1)
        self.lin2 = Linear(hidden, 1)
        self.node_type_embedding = torch.nn.Embedding(100, 32)
        self.edge_type_embedding = torch.nn.Embedding(100, 32)

    def forward(self, data):
        x, edge_index, batch = self.node_type_embedding(data.x), data.edge_index, data.batch
        z = self.z_embedding(bag_embed(self.z_initial.weight, data.pos_index, data.pos_enc, data.pos_batch))
        z = torch.cat((z, self.edge_type_embedding(data.edge_attr)), dim=-1)
        x = self.conv1(x, edge_index, z)
        xs = [x]
        for conv in self.convs:
            x = conv(x, edge_index, z)
            xs += [x]
        x = global_add_pool(torch.cat(xs, dim=1), batch)
        x = self.lin1(x)
        if x.size(0) > 1:
            x = self.bn_lin1(x)
        x = F.elu(F.dropout(x, p=self.dropout, training=self.training))
        return self.lin2(x)


class NestedGINEffQM9(torch.nn.Module):
    """qm9_models.py:25-139 (hidden 256, dropout 0 hard-coded there): continuous node features + 3-D positions, a
    node-type embedding ADDED to them, bond one-hot + distance as continuous edge attributes, mean-pool readout."""
    def __init__(self, num_layers, num_features, edge_attr_dim=5, hidden=256, dropout=0.0):
        super().__init__()
        self.dropout = dropout
        self.z_initial = torch.nn.Embedding(1800, hidden)
        self.z_embedding = _z_embedding(hidden, dropout, ReLU)
        input_dim = num_features + 3
        self.conv1 = GINEConv(_mlp(input_dim, hidden, dropout, ReLU), train_eps=True, edge_dim=hidden + edge_attr_dim)
        self.convs = torch.nn.ModuleList(
            [GINEConv(_mlp(hidden, hidden, dropout, ReLU), train_eps=True, edge_dim=hidden + edge_attr_dim)
             for _ in range(num_layers - 1)])
        self.lin1 = Linear(num_layers * hidden, hidden)
        self.bn_lin1 = BN(hidden, eps=1e-5, momentum=0.1)
        self.lin2 = Linear(hidden, 1)
        self.node_type_embedding = torch.nn.Embedding(5, input_dim)

    def forward(self, data):
        x, edge_index, batch = torch.cat([data.x, data.pos], 1), data.edge_index, data.batch
        x = x + self.node_type_embedding(data.node_type)
        z = self.z_embedding(bag_embed(self.z_initial.weight, data.pos_index, data.pos_enc, data.pos_batch))
        z = torch.cat((z, data.edge_attr), dim=-1)
        x = self.conv1(x, edge_index, z)
        xs = [x]
        for conv in self.convs:
            x = conv(x, edge_index, z)
            xs += [x]
        x = global_mean_pool(torch.cat(xs, dim=1), batch)
        x = self.lin1(x)
        if x.size(0) > 1:
            x = self.bn_lin1(x)
        x = F.relu(F.dropout(x, p=self.dropout, training=self.training))
        return self.lin2(x).view(-1)


def distance_transform(edge_index, pos, edge_attr, norm=True, max_value=None, squared=False):
    """distance.py:28-42: Euclidean length of every edge (loops -> 0), divided by the graph's maximum, appended as
    the last edge-attribute column."""
    row, col = edge_index
    d = pos[col] - pos[row]
    dist = (d ** 2).sum(1).view(-1, 1) if squared else torch.norm(d, p=2, dim=-1).view(-1, 1)
    if norm and dist.numel() > 0:
        dist = dist / (dist.max() if max_value is None else max_value)
    if edge_attr is None:
        return dist
    pseudo = edge_attr.view(-1, 1) if edge_attr.dim() == 1 else edge_attr
    return torch.cat([pseudo, dist.type_as(pseudo)], dim=-1)


class _SumEmbedding(torch.nn.Module):
    def __init__(self, dims, emb_dim, list_name):
        super().__init__()
        lst = torch.nn.ModuleList()
        for d in dims:
            emb = torch.nn.Embedding(d, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            lst.append(emb)
        setattr(self, list_name, lst)
        self._name = list_name

    def forward(self, x):
        out = 0
        for i, emb in enumerate(getattr(self, self._name)):
            out = out + emb(x[:, i])
        return out


class GINConvEff(torch.nn.Module):
    """ogb_mol_gnn.py:323-358."""
    def __init__(self, emb_dim, bond_dims=BOND_DIMS):
        super().__init__()
        self.mlp = Sequential(Linear(emb_dim, 2 * emb_dim), BN(2 * emb_dim), ReLU(), Linear(2 * emb_dim, emb_dim))
        self.eps = torch.nn.Parameter(torch.Tensor([0]))
        self.edge_encoder = _SumEmbedding(bond_dims, emb_dim, 'bond_embedding_list')
        self.edge_encoder_pos = Linear(emb_dim, emb_dim)

    def forward(self, x, edge_index, edge_attr, edge_pos):
        e = self.edge_encoder(edge_attr) + self.edge_encoder_pos(edge_pos)
        msg = F.relu(x[edge_index[0]] + e)
        agg = torch.zeros_like(x).index_add_(0, edge_index[1], msg)
        return self.mlp((1 + self.eps) * x + agg)


class GNNNodeEfficient(torch.nn.Module):
    """ogb_mol_gnn.py:614-792 with JK='last', no center pooling, no RNI."""
    def __init__(self, num_layer, emb_dim, drop_ratio=0.5, residual=False, virtual_node=True,
                 atom_dims=ATOM_DIMS, bond_dims=BOND_DIMS):
        super().__init__()
        self.num_layer, self.drop_ratio, self.residual, self.virtual_node = num_layer, drop_ratio, residual, virtual_node
        self.z_initial = torch.nn.Embedding(1800, emb_dim)
        self.z_embedding = _z_embedding(emb_dim, drop_ratio, ReLU)
        self.node_encoder = _SumEmbedding(atom_dims, emb_dim, 'atom_embedding_list')
        if virtual_node:
            self.virtualnode_embedding = torch.nn.Embedding(1, emb_dim)
            torch.nn.init.constant_(self.virtualnode_embedding.weight.data, 0)
        self.convs = torch.nn.ModuleList([GINConvEff(emb_dim, bond_dims) for _ in range(num_layer)])
        self.batch_norms = torch.nn.ModuleList([BN(emb_dim) for _ in range(num_layer)])
        if virtual_node:
            self.mlp_virtualnode_list = torch.nn.ModuleList([
                Sequential(Linear(emb_dim, 2 * emb_dim), BN(2 * emb_dim), ReLU(), Linear(2 * emb_dim, emb_dim),
                           BN(emb_dim), ReLU()) for _ in range(num_layer - 1)])

    def forward(self, data):
        x, edge_index, edge_attr, batch = data.x, data.edge_index, data.edge_attr, data.batch
        if self.virtual_node:
            vn = self.virtualnode_embedding(torch.zeros(int(batch[-1]) + 1, dtype=edge_index.dtype, device=edge_index.device))
        h_list = [self.node_encoder(x)]
        z = self.z_embedding(bag_embed(self.z_initial.weight, data.pos_index, data.pos_enc, data.pos_batch))
        for layer in range(self.num_layer):
            if self.virtual_node:
                h_list[layer] = h_list[layer] + vn[batch]
            h = self.batch_norms[layer](self.convs[layer](h_list[layer], edge_index, edge_attr, z))
            if layer == self.num_layer - 1:
                h = F.dropout(h, self.drop_ratio, training=self.training)
            else:
                h = F.dropout(F.relu(h), self.drop_ratio, training=self.training)
            if self.residual:
                h = h + h_list[layer]
            h_list.append(h)
            if self.virtual_node and layer < self.num_layer - 1:
                tmp = global_add_pool(h_list[layer], batch) + vn
                upd = F.dropout(self.mlp_virtualnode_list[layer](tmp), self.drop_ratio, training=self.training)
                vn = vn + upd if self.residual else upd
        return h_list[-1]


class GNNOgbEff(torch.nn.Module):
    """ogb_mol_gnn.py:66-117,252-261: GNN(gnn_type='gin_eff', JK='last', graph_pooling='mean')."""
    def __init__(self, num_tasks, num_layer=5, emb_dim=300, virtual_node=True, residual=False, drop_ratio=0.5,
                 atom_dims=ATOM_DIMS, bond_dims=BOND_DIMS):
        super().__init__()
        self.gnn_node = GNNNodeEfficient(num_layer, emb_dim, drop_ratio, residual, virtual_node, atom_dims, bond_dims)
        self.graph_pred_linear = Linear(emb_dim, num_tasks)

    def forward(self, data):
        return self.graph_pred_linear(global_mean_pool(self.gnn_node(data), data.batch))


# ---- collation (batch.py:25-149), only the keys the efficient path carries ---------------------------------------
class RefBatch(object):
    pass


def collate(graphs):
    """graphs: list of dicts with x, edge_index, [edge_attr], y, pos_enc, pos_index, pos_batch (torch tensors).
    Rules: edge_index += cumulative num_nodes (:112-113), pos_batch += cumulative (max+1) (:70-71),
    pos_enc / pos_index concatenated unchanged (:72-73), batch vector (:120-123)."""
    b = RefBatch()
    node_off, pb_off = 0, 0
    cols = {k: [] for k in ('x', 'edge_index', 'edge_attr', 'y', 'pos_enc', 'pos_index', 'pos_batch', 'batch', 'pos',
                            'node_type')}
    for i, g in enumerate(graphs):
        n = g['x'].size(0)
        cols['x'].append(g['x'])
        cols['edge_index'].append(g['edge_index'] + node_off)
        if g.get('edge_attr') is not None:
            cols['edge_attr'].append(g['edge_attr'])
        cols['y'].append(g['y'].view(-1) if g['y'].dim() == 0 else g['y'])
        cols['pos_enc'].append(g['pos_enc'])
        cols['pos_index'].append(g['pos_index'])
        cols['pos_batch'].append(g['pos_batch'] + pb_off)
        cols['batch'].append(torch.full((n, ), i, dtype=torch.long))
        for k in ('pos', 'node_type'):            # QM9 extras: plain concatenation (batch.py default rule)
            if g.get(k) is not None:
                cols[k].append(g[k])
        node_off += n
        pb_off += int(g['pos_batch'].max()) + 1
    b.x = torch.cat(cols['x'], 0)
    b.edge_index = torch.cat(cols['edge_index'], 1)
    b.edge_attr = torch.cat(cols['edge_attr'], 0) if cols['edge_attr'] else None
    b.y = torch.cat(cols['y'], 0)
    b.pos_enc, b.pos_index, b.pos_batch = (torch.cat(cols[k], 0) for k in ('pos_enc', 'pos_index', 'pos_batch'))
    b.batch = torch.cat(cols['batch'], 0)
    for k in ('pos', 'node_type'):
        if cols[k]:
            setattr(b, k, torch.cat(cols[k], 0))
    b.num_graphs = len(graphs)
    return b
