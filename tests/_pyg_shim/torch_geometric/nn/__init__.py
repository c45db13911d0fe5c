"""GINEConv / pooling stand-ins with PyG 2.0.4 semantics.

GINEConv follows the in-tree restatement at
/root/reference/GraphGPS/graphgps/layer/gine_conv_layer.py:18-35,49-84 (minus the LapPE scaling).
"""
import torch


def _reset(module):
    if hasattr(module, 'reset_parameters'):
        module.reset_parameters()
    elif hasattr(module, 'children'):
        for m in module.children():
            _reset(m)


def global_add_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    out = x.new_zeros((size, ) + tuple(x.shape[1:]))
    return out.index_add_(0, batch, x)


def global_mean_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    out = global_add_pool(x, batch, size)
    cnt = torch.zeros(size, dtype=x.dtype, device=x.device).index_add_(
        0, batch, torch.ones_like(batch, dtype=x.dtype)).clamp_(min=1)
    return out / cnt.view(-1, *([1] * (x.dim() - 1)))


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr='add', **kwargs):
        super().__init__()
        assert aggr == 'add'

    def propagate(self, edge_index, x=None, edge_attr=None, size=None):
        msg = self.message(x[edge_index[0]], edge_attr)
        out = torch.zeros((x.size(0), msg.size(1)), dtype=msg.dtype, device=msg.device)
        out.index_add_(0, edge_index[1], msg)
        return self.update(out)

    def update(self, aggr_out):
        return aggr_out


class GINEConv(MessagePassing):
    def __init__(self, nn, eps=0., train_eps=False, edge_dim=None, **kwargs):
        super().__init__(aggr='add')
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer('eps', torch.Tensor([eps]))
        if edge_dim is not None:
            in_channels = nn[0].in_features
            self.lin = torch.nn.Linear(edge_dim, in_channels)
        else:
            self.lin = None
        self.reset_parameters()

    def reset_parameters(self):
        _reset(self.nn)
        self.eps.data.fill_(self.initial_eps)
        if self.lin is not None:
            self.lin.reset_parameters()

    def forward(self, x, edge_index, edge_attr=None, size=None):
        out = self.propagate(edge_index, x=x, edge_attr=edge_attr)
        out = out + (1 + self.eps) * x
        return self.nn(out)

    def message(self, x_j, edge_attr):
        if self.lin is not None:
            edge_attr = self.lin(edge_attr)
        return (x_j + edge_attr).relu()


class GINConv(MessagePassing):
    pass


class GCNConv(MessagePassing):
    pass


class GATConv(MessagePassing):
    pass
