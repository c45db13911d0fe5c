"""The hot-path operators registered with the PyTorch dispatcher (`torch.library`): `torch.ops.escgnn.*`.

SURVEY.md section 8(b) lists the operators a replacement must export "registered with TORCH_LIBRARY + autograd Functions";
the product's binding is ctypes over the C-ABI (`_lib.py`), so the registration is done from Python with `torch.library.custom_op`
(the same dispatcher entries a C++ `TORCH_LIBRARY` block creates: schema, CUDA implementation, fake / meta implementation for
tracing, autograd formula).  Every operator takes plain tensors (the CSR pieces of `ops.GraphIndex` are passed explicitly) and runs
the same sm_100a kernels as the drop-in modules; there is no CPU implementation -- calling them with CPU tensors raises.

    torch.ops.escgnn.bag_embed(weight, pos_index, pos_enc, rec_ptr, n_edges)               M1   run_graphcount.py:155
    torch.ops.escgnn.gine_aggregate(x, edge_feat, eps, src, dst, dst_ptr, dst_perm, src_ptr, src_perm)   M3   gine_conv_layer.py:56-84
    torch.ops.escgnn.segment_pool(x, ptr, segments, mean)                                  M4   global_add_pool / global_mean_pool
    torch.ops.escgnn.linear(x, weight, bias)                                               M2 / M3  nn.Linear on the tcgen05 3xTF32 GEMM
    torch.ops.escgnn.edge_distance(pos, edge_index, squared)                               D1   distance.py:29-37
"""
import ctypes
import types
from typing import Optional

import torch
from torch import Tensor

from . import _lib, ops

_p = ops._p


def _idx(src, dst, dst_ptr, dst_perm, src_ptr, src_perm, n_nodes):
    return types.SimpleNamespace(src=src, dst=dst, dst_ptr=dst_ptr, dst_perm=dst_perm, src_ptr=src_ptr, src_perm=src_perm,
                                 num_edges=int(src.numel()), num_nodes=int(n_nodes))


# ---------------------------------------------------------------------------------------------------- M1 bag-embed
@torch.library.custom_op('escgnn::bag_embed', mutates_args=(), device_types='cuda')
def bag_embed(weight: Tensor, pos_index: Tensor, pos_enc: Tensor, rec_ptr: Tensor, n_edges: int) -> Tensor:
    return ops._BagEmbed.forward(types.SimpleNamespace(save_for_backward=lambda *a: None), weight, pos_index.contiguous(),
                                 pos_enc.contiguous(), rec_ptr, n_edges)


@bag_embed.register_fake
def _(weight, pos_index, pos_enc, rec_ptr, n_edges):
    return weight.new_empty((n_edges, weight.size(1)))


@torch.library.custom_op('escgnn::bag_embed_backward', mutates_args=(), device_types='cuda')
def bag_embed_backward(grad: Tensor, pos_index: Tensor, pos_enc: Tensor, rec_ptr: Tensor, rows: int) -> Tensor:
    g = grad.contiguous()
    dW = torch.zeros((rows, g.size(1)), dtype=torch.float32, device=g.device)
    _lib.check(_lib.lib().escgnn_bag_embed_bwd(_p(g), g.size(1), _p(pos_index), _p(pos_enc), _p(rec_ptr), None, None, None, g.size(0),
                                               _p(dW), None, ops._stream(g)), 'bag_embed_bwd')
    return dW


@bag_embed_backward.register_fake
def _(grad, pos_index, pos_enc, rec_ptr, rows):
    return grad.new_empty((rows, grad.size(1)))


def _bag_setup(ctx, inputs, output):
    weight, pos_index, pos_enc, rec_ptr, _ = inputs
    ctx.save_for_backward(pos_index, pos_enc, rec_ptr)
    ctx.rows = weight.size(0)


def _bag_backward(ctx, g):
    pos_index, pos_enc, rec_ptr = ctx.saved_tensors
    return torch.ops.escgnn.bag_embed_backward(g, pos_index, pos_enc, rec_ptr, ctx.rows), None, None, None, None


bag_embed.register_autograd(_bag_backward, setup_context=_bag_setup)


# ---------------------------------------------------------------------------------------------------- M3 GINE aggregation
@torch.library.custom_op('escgnn::gine_aggregate', mutates_args=(), device_types='cuda')
def gine_aggregate(x: Tensor, edge_feat: Tensor, eps: Tensor, src: Tensor, dst: Tensor, dst_ptr: Tensor, dst_perm: Tensor,
                   src_ptr: Tensor, src_perm: Tensor) -> Tensor:
    ctx = types.SimpleNamespace(save_for_backward=lambda *a: None)
    return ops._GineAggregate.forward(ctx, x, edge_feat, eps, _idx(src, dst, dst_ptr, dst_perm, src_ptr, src_perm, x.size(0)))


@gine_aggregate.register_fake
def _(x, edge_feat, eps, src, dst, dst_ptr, dst_perm, src_ptr, src_perm):
    return torch.empty_like(x)


@torch.library.custom_op('escgnn::gine_aggregate_backward', mutates_args=(), device_types='cuda')
def gine_aggregate_backward(grad: Tensor, x: Tensor, edge_feat: Tensor, eps: Tensor, dst: Tensor, src_ptr: Tensor,
                            src_perm: Tensor) -> tuple[Tensor, Tensor, Tensor]:
    g, x, edge_feat = grad.contiguous(), x.contiguous(), edge_feat.contiguous()
    N, C = x.shape
    gx, ge = torch.empty_like(x), torch.empty_like(edge_feat)
    dots = torch.empty(N, dtype=torch.float32, device=x.device)
    geps = torch.empty(1, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().escgnn_gine_aggregate_bwd(_p(g), _p(x), _p(edge_feat), _p(dst), _p(src_ptr), _p(src_perm), _p(eps), N, C,
                                                    _p(gx), _p(ge), _p(dots), _p(geps), None, ops._stream(x)), 'gine_aggregate_bwd')
    return gx, ge, geps


@gine_aggregate_backward.register_fake
def _(grad, x, edge_feat, eps, dst, src_ptr, src_perm):
    return torch.empty_like(x), torch.empty_like(edge_feat), eps.new_empty((1, ))


def _gine_setup(ctx, inputs, output):
    x, edge_feat, eps, src, dst, dst_ptr, dst_perm, src_ptr, src_perm = inputs
    ctx.save_for_backward(x, edge_feat, eps, dst, src_ptr, src_perm)


def _gine_backward(ctx, g):
    x, edge_feat, eps, dst, src_ptr, src_perm = ctx.saved_tensors
    gx, ge, geps = torch.ops.escgnn.gine_aggregate_backward(g, x, edge_feat, eps, dst, src_ptr, src_perm)
    return gx, ge, geps.view_as(eps), None, None, None, None, None, None


gine_aggregate.register_autograd(_gine_backward, setup_context=_gine_setup)


# ---------------------------------------------------------------------------------------------------- M4 pooling
@torch.library.custom_op('escgnn::segment_pool', mutates_args=(), device_types='cuda')
def segment_pool(x: Tensor, ptr: Tensor, segments: int, mean: bool) -> Tensor:
    ctx = types.SimpleNamespace(save_for_backward=lambda *a: None)
    return ops._SegmentPool.forward(ctx, x, ptr, segments, mean)


@segment_pool.register_fake
def _(x, ptr, segments, mean):
    return x.new_empty((segments, x.size(1)))


@torch.library.custom_op('escgnn::segment_pool_backward', mutates_args=(), device_types='cuda')
def segment_pool_backward(grad: Tensor, ptr: Tensor, rows: int, mean: bool) -> Tensor:
    g = grad.contiguous()
    gx = torch.zeros((rows, g.size(1)), dtype=torch.float32, device=g.device)
    _lib.check(_lib.lib().escgnn_segment_pool_bwd(_p(g), _p(ptr), g.size(0), g.size(1), int(mean), _p(gx), ops._stream(g)),
               'segment_pool_bwd')
    return gx


@segment_pool_backward.register_fake
def _(grad, ptr, rows, mean):
    return grad.new_empty((rows, grad.size(1)))


def _pool_setup(ctx, inputs, output):
    x, ptr, segments, mean = inputs
    ctx.save_for_backward(ptr)
    ctx.rows, ctx.mean = x.size(0), bool(mean)


def _pool_backward(ctx, g):
    (ptr, ) = ctx.saved_tensors
    return torch.ops.escgnn.segment_pool_backward(g, ptr, ctx.rows, ctx.mean), None, None, None


segment_pool.register_autograd(_pool_backward, setup_context=_pool_setup)


# ---------------------------------------------------------------------------------------------------- nn.Linear on tcgen05
@torch.library.custom_op('escgnn::gemm', mutates_args=(), device_types='cuda')
def gemm(a: Tensor, a_mn_major: bool, b: Tensor, b_mn_major: bool, bias: Optional[Tensor]) -> Tensor:
    """C[M, N] = A B^T (+ bias); *_mn_major: the operand is stored transposed ([K, M] / [K, N] row-major)."""
    a, b = a.contiguous(), b.contiguous()
    M = a.size(1) if a_mn_major else a.size(0)
    K = a.size(0) if a_mn_major else a.size(1)
    N = b.size(1) if b_mn_major else b.size(0)
    c = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if M and N:
        ops._gemm(a, a_mn_major, b, b_mn_major, c, bias, M, N, K)
    return c


@gemm.register_fake
def _(a, a_mn_major, b, b_mn_major, bias):
    return a.new_empty((a.size(1) if a_mn_major else a.size(0), b.size(1) if b_mn_major else b.size(0)))


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """y = x W^T + b with forward, dgrad and wgrad on the tcgen05 GEMM (autograd through `ops._LinearFn`)."""
    return ops._LinearFn.apply(x, weight, bias)


# ---------------------------------------------------------------------------------------------------- D1 edge distance
@torch.library.custom_op('escgnn::edge_distance', mutates_args=(), device_types='cuda')
def edge_distance(pos: Tensor, edge_index: Tensor, squared: bool) -> Tensor:
    """||pos[col] - pos[row]||_2 (or its square) per edge, [E, 1] (distance.py:29-37)."""
    from .ops_distance import edge_distance as _ed
    return _ed(pos, edge_index, squared=squared, norm=False)[0]


@edge_distance.register_fake
def _(pos, edge_index, squared):
    return pos.new_empty((edge_index.size(1), 1))


def index_tensors(index):
    """The tensor pieces of an `ops.GraphIndex`, in the order `torch.ops.escgnn.gine_aggregate` takes them."""
    return index.src, index.dst, index.dst_ptr, index.dst_perm, index.src_ptr, index.src_perm
