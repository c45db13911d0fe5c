"""PyG 2.0.4 util semantics used by the reference encoder."""
import numpy as np
import scipy.sparse as ssp
import torch


def maybe_num_nodes(edge_index, num_nodes=None):
    if num_nodes is not None:
        return num_nodes
    return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    if edge_attr is None:
        return edge_index, None
    return edge_index, edge_attr[mask]


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    N = maybe_num_nodes(edge_index, num_nodes)
    loop_index = torch.arange(0, N, dtype=torch.long, device=edge_index.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    if edge_attr is not None:
        loop_attr = edge_attr.new_full((N, ) + edge_attr.size()[1:], 1.)
        edge_attr = torch.cat([edge_attr, loop_attr], dim=0)
    edge_index = torch.cat([edge_index, loop_index], dim=1)
    return edge_index, edge_attr


def add_remaining_self_loops(*a, **k):  # import-time name only
    raise NotImplementedError


def to_dense_batch(*a, **k):  # import-time name only
    raise NotImplementedError


def to_dense_adj(*a, **k):
    raise NotImplementedError


def dropout_adj(*a, **k):
    raise NotImplementedError


def degree(index, num_nodes=None, dtype=None):
    N = maybe_num_nodes(index, num_nodes)
    out = torch.zeros((N, ), dtype=dtype, device=index.device)
    one = torch.ones((index.size(0), ), dtype=out.dtype, device=out.device)
    return out.scatter_add_(0, index, one)


def to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
    row, col = edge_index.cpu()
    if edge_attr is None:
        edge_attr = torch.ones(row.size(0))  # float32: this is what makes the rd block float32 (F7)
    else:
        edge_attr = edge_attr.view(-1).cpu()
    N = maybe_num_nodes(edge_index, num_nodes)
    return ssp.coo_matrix((edge_attr.numpy(), (row.numpy(), col.numpy())), (N, N))
