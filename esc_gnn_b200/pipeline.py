"""Device-resident hot path end to end: raw graphs -> structural encoding -> collation -> NestedGIN_eff train step.

This is the B200-native composition of the reference's two hot paths (the offline `pre_transform` of
run_zinc.py:141-146 / dataset_zinc.py:76-85 and the `train()` loop of run_zinc.py:266-289): the encoder's packed
records feed the model's bag-embed kernel directly, collation happens on the device, Adam is one kernel over a
flat buffer and the data-parallel exchange one NCCL all-reduce.  The drop-in (per-graph `create_subgraphs`,
`Batch.from_data_list`, `model(data)`) stays available and is tested equal to this path.
"""
import ctypes

import numpy as np
import torch

from . import _lib, synth
from .optim import FlatAdam
from .transform import encode_batch


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class DeviceBatch(object):
    """What the models read (`x, edge_index, edge_attr, y, batch, num_graphs`) plus the encoder's packed records."""
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device):
        return self

    def __contains__(self, key):
        return key in self.__dict__


class RawBatch(object):
    """Raw (un-encoded) graphs of one step, packed: graph-local edge ids + per-graph pointers + features."""
    KEYS = ('src', 'dst', 'edge_ptr', 'node_ptr', 'x', 'edge_attr', 'y')

    def __init__(self, src, dst, edge_ptr, node_ptr, x, edge_attr, y):
        self.src, self.dst, self.edge_ptr, self.node_ptr, self.x, self.edge_attr, self.y = \
            src, dst, edge_ptr, node_ptr, x, edge_attr, y
        nn = np.diff(np.asarray(node_ptr.cpu() if torch.is_tensor(node_ptr) else node_ptr))
        ee = np.diff(np.asarray(edge_ptr.cpu() if torch.is_tensor(edge_ptr) else edge_ptr))
        self.num_graphs = len(nn)
        self.max_nodes, self.max_in_edges = int(nn.max()), int(ee.max())
        self.max_loop_edges = int((nn + ee).max())
        self.num_nodes = int(nn.sum())

    @staticmethod
    def synth(config, start, count, pin=True):
        gs = [synth.make_graph(config, i) for i in range(start, start + count)]
        cat = lambda k, ax=0: np.concatenate([np.atleast_1d(g[k]) for g in gs], axis=ax)
        ei = np.concatenate([g['edge_index'] for g in gs], axis=1)
        eptr = np.cumsum([0] + [g['edge_index'].shape[1] for g in gs])
        nptr = np.cumsum([0] + [g['num_nodes'] for g in gs])
        t = lambda a: (torch.as_tensor(np.ascontiguousarray(a)).pin_memory() if pin and torch.cuda.is_available()
                       else torch.as_tensor(np.ascontiguousarray(a)))
        ea = t(cat('edge_attr')) if 'edge_attr' in gs[0] else None
        return RawBatch(t(ei[0]), t(ei[1]), t(eptr.astype(np.int64)), t(nptr.astype(np.int64)), t(cat('x')), ea,
                        t(cat('y').astype(np.float32)))

    def h2d_bytes(self):
        return sum(getattr(self, k).numel() * getattr(self, k).element_size() for k in self.KEYS
                   if getattr(self, k) is not None)

    def cuda(self, non_blocking=True):
        out = RawBatch.__new__(RawBatch)
        out.__dict__.update(self.__dict__)
        for k in self.KEYS:
            v = getattr(self, k)
            setattr(out, k, v.cuda(non_blocking=non_blocking) if v is not None else None)
        out.edge_ptr_host, out.node_ptr_host = self.edge_ptr, self.node_ptr
        return out


def encode_and_collate(raw, h, use_rd, self_loop, timings=None):
    """Device: E1-E5 encoding of a RawBatch (already on the GPU) + collation rules of batch.py:52-123."""
    L = _lib.lib()
    max_edges = raw.max_loop_edges if self_loop else raw.max_in_edges
    r = encode_batch(raw.src, raw.dst, raw.edge_ptr, raw.node_ptr, h, use_rd, self_loop, expand=False,
                     max_nodes=raw.max_nodes, max_edges=max_edges, n_total=raw.num_nodes, timings=timings)
    E, dev = r.num_edges, raw.src.device
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ei = torch.empty((2, E), dtype=torch.int64, device=dev)
    _lib.check(L.escgnn_collate_edges(_p(r.edge_index[0]), _p(r.edge_index[1]), _p(r.edge_graph), _p(raw.node_ptr), E,
                                      _p(ei[0]), _p(ei[1]), None, st), 'collate_edges')
    batch = torch.empty(raw.num_nodes, dtype=torch.int64, device=dev)
    _lib.check(L.escgnn_ptr_to_ids(_p(raw.node_ptr), raw.num_graphs, raw.num_nodes, _p(batch), None, st), 'ptr_to_ids')
    edge_attr = raw.edge_attr
    if self_loop and edge_attr is not None:          # E1 on attributes: synthetic graphs carry no loops
        tail = edge_attr.new_full((raw.num_nodes, ) + tuple(edge_attr.shape[1:]), 1)
        edge_attr = _interleave_loop_attr(edge_attr, tail, raw, r)
    return DeviceBatch(x=raw.x, edge_index=ei, edge_attr=edge_attr, y=raw.y, batch=batch, num_graphs=raw.num_graphs,
                       rec=r.rec, rec_off=r.rec_off, rec_nnz=r.rec_nnz, nnz=r.nnz, num_edges=E)


def _interleave_loop_attr(edge_attr, tail, raw, r):
    """Per graph: [its edge attrs..., its N loop rows of ones] (add_self_loops appends loops after each graph's edges)."""
    E_in = edge_attr.size(0)
    eptr, nptr = raw.edge_ptr, raw.node_ptr
    out = edge_attr.new_empty((E_in + tail.size(0), ) + tuple(edge_attr.shape[1:]))
    g_of_e = torch.bucketize(torch.arange(E_in, device=edge_attr.device), eptr[1:], right=True)
    out[torch.arange(E_in, device=edge_attr.device) + nptr[g_of_e]] = edge_attr
    g_of_n = torch.bucketize(torch.arange(tail.size(0), device=edge_attr.device), nptr[1:], right=True)
    out[torch.arange(tail.size(0), device=edge_attr.device) + eptr[g_of_n + 1]] = tail
    return out


class TrainPipeline(object):
    """encode -> collate -> forward -> loss -> backward -> (all-reduce) -> Adam, for one config."""

    def __init__(self, model, loss_fn, h, use_rd, self_loop, lr=1e-3, distributed=False):
        self.model, self.loss_fn = model, loss_fn
        self.h, self.use_rd, self.self_loop = h, use_rd, self_loop
        self.opt = FlatAdam(model.parameters(), lr=lr)
        self.distributed = distributed

    def step_device(self, raw_dev, timings=None):
        """`raw_dev`: RawBatch already resident in HBM. Returns the loss as a device scalar (no sync)."""
        batch = encode_and_collate(raw_dev, self.h, self.use_rd, self.self_loop, timings=timings)
        self.opt.zero_grad()
        loss = self.loss_fn(self.model(batch), batch.y)
        loss.backward()
        world = self.opt.all_reduce_grads() if self.distributed else 1
        self.opt.step(world)
        return loss

    def step_host(self, raw_host, timings=None):
        """`raw_host`: RawBatch in pinned host memory. H2D of the raw graphs and D2H of the loss are inside."""
        loss = self.step_device(raw_host.cuda(non_blocking=True), timings=timings)
        return float(loss.item())
