"""Binding of the `edge_distance` kernel (K2). CPU tensors are staged through the GPU; there is no CPU fallback."""
import ctypes

import torch

from . import _lib


def edge_distance(pos, edge_index, squared=False, norm=True, max_value=None):
    """Returns (dist [E,1], rel [E,dim]) on the device/dtype of `pos` (distance.py:29-47)."""
    if not torch.cuda.is_available():
        raise RuntimeError('esc_gnn_b200: no CUDA device -- Distance has no CPU fallback')
    home = pos.device
    dev = home if pos.is_cuda else torch.device('cuda', torch.cuda.current_device())
    p32 = pos.to(dev, torch.float32).contiguous()
    ei = edge_index.to(dev, torch.int64).contiguous()
    E, dim = ei.size(1), p32.size(1)
    dist = torch.empty((E, 1), dtype=torch.float32, device=dev)
    rel = torch.empty((E, dim), dtype=torch.float32, device=dev)
    scratch = torch.zeros(1, dtype=torch.int32, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().escgnn_edge_distance(P(p32), dim, P(ei[0]), P(ei[1]), E, int(bool(squared)), int(bool(norm)),
                                               ctypes.c_float(0.0 if max_value is None else float(max_value)),
                                               P(dist), P(rel), P(scratch), st), 'edge_distance')
    return dist.to(home, pos.dtype), rel.to(home, pos.dtype)
