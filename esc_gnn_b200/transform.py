"""Drop-in for the reference's efficient structural-encoding transform, backed by the sm_100a kernels.

`create_subgraphs` keeps the signature and output contract of
/root/reference/utils_edge_efficient.py:20-152 (same keyword names, same `Data` fields `pos_enc`, `pos_index`,
`pos_batch`, same self-loop rewrite of `edge_index` / `edge_attr`, same exceptions for the degenerate inputs).
`encode_batch` is the batched entry the kernels really want: many graphs per launch, results left on the device
in compact form for the model's bag-embed kernel, or expanded to the reference's int64 triple.

No CPU fallback: without libescgnn_b200.so / a CUDA device these functions raise.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

_ctx_cache = {}
_host_encoders = {}


def _ctx(device_index):
    c = _ctx_cache.get(device_index)
    if c is None:
        if not torch.cuda.is_available():
            raise RuntimeError('esc_gnn_b200: no CUDA device -- the encoder has no CPU fallback')
        c = _lib.lib().escgnn_ctx_create(int(device_index))
        if not c:
            raise RuntimeError('esc_gnn_b200: escgnn_ctx_create failed on device %d' % device_index)
        _ctx_cache[device_index] = c
    return c


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class EncodedBatch(object):
    """Result of `encode_batch`.

    Reference-contract fields (int64): `edge_index` [2, E_out] (graph-local ids, after the self-loop rewrite),
    `edge_ptr` [G+1], and -- when expanded -- `pos_enc`, `pos_index`, `pos_batch` [nnz].
    Compact device fields (what the model consumes directly): `rec` uint32 [nnz] = index | count << 11,
    `rec_off` int64 [E_out], `rec_nnz` int32 [E_out].
    """
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _np_view(ptr, n, dtype):
    """numpy view of `n` items of pinned host memory owned by the encoder context (no copy)."""
    if n == 0 or not ptr:
        return np.empty(0, dtype=dtype)
    ctype = {np.uint32: ctypes.c_uint32, np.int32: ctypes.c_int32, np.int64: ctypes.c_int64}[dtype]
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(n, ))


class HostEncodedBatch(object):
    """Result of the host front end in COMPACT form (what crosses PCIe): `rec` uint32 [nnz] = index | count << 11, ascending
    inside an edge, the records of edge e at rec[rec_off[e] : rec_off[e] + rec_nnz[e]]; `edge_index` int64 [2, E_out] and
    `edge_ptr` int64 [G+1] after the self-loop rewrite.  The arrays are views of pinned memory owned by the encoder context
    and stay valid until the same slot is submitted again (`detach()` copies them out).

    The reference's int64 triple (`pos_enc`, `pos_index`, `pos_batch`, utils_edge_efficient.py:139-151) is produced on
    first access, on the host, by all cores (`expand()`)."""

    def __init__(self, rec, rec_off, rec_nnz, eo_src, eo_dst, edge_ptr, num_edges, nnz, local_ordinals):
        self.rec, self.rec_off, self.rec_nnz = rec, rec_off, rec_nnz
        self._eo_src, self._eo_dst, self._edge_ptr = eo_src, eo_dst, edge_ptr
        self.num_edges, self.nnz, self.local_ordinals = num_edges, nnz, local_ordinals
        self._triple = None

    @property
    def edge_index(self):
        return torch.from_numpy(np.stack([self._eo_src, self._eo_dst]))

    @property
    def edge_ptr(self):
        return torch.from_numpy(np.array(self._edge_ptr))

    def detach(self):
        """Copy the compact arrays out of the context's pinned arena (they are overwritten by the slot's next run)."""
        for k in ('rec', 'rec_off', 'rec_nnz', '_eo_src', '_eo_dst', '_edge_ptr'):
            setattr(self, k, np.array(getattr(self, k)))
        return self

    def expand(self, out=None, threads=None):
        """(pos_enc, pos_index, pos_batch) as int64 CPU tensors; `out` = dict of preallocated (e.g. pinned) tensors."""
        if self._triple is None:
            K = self.nnz
            if out is not None:
                pe, pi, pb = out['pos_enc'][:K], out['pos_index'][:K], out['pos_batch'][:K]
            else:
                pe, pi, pb = (torch.empty(K, dtype=torch.int64) for _ in range(3))
            ep = np.ascontiguousarray(self._edge_ptr)
            p = lambda a: ctypes.c_void_p(a.ctypes.data)
            _lib.check(_lib.lib().escgnn_expand_records_host(p(self.rec), p(self.rec_off), p(self.rec_nnz), p(ep), len(ep) - 1,
                                                             int(self.local_ordinals), _ptr(pe), _ptr(pi), _ptr(pb),
                                                             int(threads or (1 if K < (1 << 16) else (os.cpu_count() or 1)))),
                       'expand_records_host')        # (a single graph is not worth waking a thread team for)
            self._triple = (pe, pi, pb)
        return self._triple

    pos_enc = property(lambda self: self.expand()[0])
    pos_index = property(lambda self: self.expand()[1])
    pos_batch = property(lambda self: self.expand()[2])


class HostEncoder(object):
    """Pipelined host front end over the two slots of an encoder context (C-ABI escgnn_encode_host_submit / _wait).

        enc = HostEncoder(h=3, use_rd=True)
        for result in enc.stream(chunks):        # chunks: iterable of (src, dst, edge_ptr, node_ptr) int64 host arrays
            ...                                  # HostEncodedBatch, valid until two more chunks have been submitted

    While the caller consumes chunk k, the kernels of chunk k+1 are already running and its own D2H ran under them."""

    def __init__(self, h, use_rd=False, self_loop=False, local_ordinals=False, device=None):
        self.h, self.use_rd, self.self_loop, self.local_ordinals = int(h), bool(use_rd), bool(self_loop), bool(local_ordinals)
        if device is None:          # the calling thread's current CUDA device
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = int(device)
        self.ctx = _ctx(self.device)
        self.L = _lib.lib()

    def submit(self, slot, src, dst, edge_ptr, node_ptr):
        a = [np.ascontiguousarray(x.numpy() if torch.is_tensor(x) else x, dtype=np.int64) for x in (src, dst, edge_ptr, node_ptr)]
        G = a[2].shape[0] - 1
        p = lambda x: ctypes.c_void_p(x.ctypes.data)
        _lib.check(self.L.escgnn_encode_host_submit(self.ctx, slot, p(a[0]), p(a[1]), p(a[2]), p(a[3]), G, self.h,
                                                    int(self.use_rd), int(self.self_loop)), 'encode_host_submit')
        return G

    def wait(self, slot, G):
        e_out, nnz, bits = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_uint32(0)
        ptrs = [ctypes.c_void_p() for _ in range(6)]
        rc = self.L.escgnn_encode_host_wait(self.ctx, slot, ctypes.byref(e_out), ctypes.byref(nnz), ctypes.byref(bits),
                                            *[ctypes.byref(q) for q in ptrs])
        if rc == -4:
            _lib.raise_data_errors(bits.value)
        _lib.check(rc, 'encode_host_wait')
        E, K = e_out.value, nnz.value
        v = [q.value for q in ptrs]
        return HostEncodedBatch(_np_view(v[0], K, np.uint32), _np_view(v[1], E, np.int64), _np_view(v[2], E, np.int32),
                                _np_view(v[3], E, np.int64), _np_view(v[4], E, np.int64), _np_view(v[5], G + 1, np.int64), E, K,
                                self.local_ordinals)

    def encode(self, src, dst, edge_ptr, node_ptr):
        """One chunk, synchronously (slot 0)."""
        return self.wait(0, self.submit(0, src, dst, edge_ptr, node_ptr))

    def stream(self, chunks):
        pending = None
        k = 0
        for chunk in chunks:
            slot = k & 1
            G = self.submit(slot, *chunk)
            if pending is not None:
                yield self.wait(*pending)
            pending = (slot, G)
            k += 1
        if pending is not None:
            yield self.wait(*pending)


def encode_batch_host(src, dst, edge_ptr, node_ptr, h, use_rd=False, self_loop=False, local_ordinals=False,
                      device=None, out=None, compact=False):
    """HOST buffers in, HOST buffers out through the C-ABI host front end (H2D + kernels + D2H inside the call).

    src/dst/edge_ptr/node_ptr: int64 numpy arrays or CPU tensors (graph-local node ids).  Only the compact records cross
    PCIe (4 bytes per record, pinned staging on both sides); `compact=True` returns them as a HostEncodedBatch (views of the
    context's pinned arena, the triple expanded lazily), otherwise the reference's int64 triple is expanded on the host
    (all cores) into `out` (dict of preallocated tensors, e.g. pinned) or fresh tensors and an EncodedBatch of CPU
    tensors is returned."""
    G = len(edge_ptr) - 1
    if G == 0:
        z = lambda *shape: torch.zeros(shape, dtype=torch.int64)
        return EncodedBatch(edge_index=z(2, 0), edge_ptr=z(1), pos_enc=z(0), pos_index=z(0), pos_batch=z(0), num_edges=0, nnz=0)
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    key = (int(h), bool(use_rd), bool(self_loop), bool(local_ordinals), int(device))
    enc = _host_encoders.get(key)
    if enc is None:
        enc = _host_encoders[key] = HostEncoder(h, use_rd, self_loop, local_ordinals, device)
    r = enc.encode(src, dst, edge_ptr, node_ptr)
    if compact:
        return r
    E, K = r.num_edges, r.nnz
    pe, pi, pb = r.expand(out)
    if out is not None:
        ei, eptr = out['edge_index'][:, :E], out['edge_ptr'][:G + 1]
        ei[0].copy_(torch.from_numpy(r._eo_src)); ei[1].copy_(torch.from_numpy(r._eo_dst))
        eptr.copy_(torch.from_numpy(r._edge_ptr))
    else:
        ei, eptr = r.edge_index, r.edge_ptr
    return EncodedBatch(edge_index=ei, edge_ptr=eptr, pos_enc=pe, pos_index=pi, pos_batch=pb, num_edges=E, nnz=K)


class _Timer(object):
    """Optional CUDA-event brackets around the kernels of one encode_batch call (bench / profiling only)."""
    def __init__(self, sink):
        self.sink = sink

    def __call__(self, name):
        return _Span(self.sink, name)


class _Span(object):
    def __init__(self, sink, name):
        self.sink, self.name = sink, name

    def __enter__(self):
        if self.sink is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.sink is not None:
            self.b.record()
            self.sink.setdefault(self.name, []).append((self.a, self.b))


def encode_batch(src, dst, edge_ptr, node_ptr, h, use_rd=False, self_loop=False, expand=True, local_ordinals=False,
                 max_nodes=None, max_edges=None, n_total=None, timings=None):
    """DEVICE tensors in, DEVICE tensors out, on the current torch stream.

    src/dst: int64 CUDA tensors [E_in] (graph-local ids); edge_ptr/node_ptr: int64 [G+1], CPU or CUDA
    (`max_nodes` / `max_edges` must be given when they are CUDA tensors, so no sync is needed to size the launch).
    One device->host read of two counters (total records, error bits) sizes the output."""
    L = _lib.lib()
    if not src.is_cuda:
        raise RuntimeError('encode_batch needs CUDA tensors (use encode_batch_host for CPU buffers); no CPU fallback')
    dev = src.device
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    G = edge_ptr.numel() - 1
    classes = None
    if max_nodes is None or max_edges is None:
        ep, npt = edge_ptr.cpu(), node_ptr.cpu()
        nn = (npt[1:] - npt[:-1])
        ee = (ep[1:] - ep[:-1])
        eo_n = (ee + nn) if self_loop else ee
        max_nodes = int(nn.max()) if G else 0
        max_edges = int(eo_n.max()) if G else 0
        n_total = int(npt[-1])
        if G and max_nodes > 96:
            # size classes by on-chip footprint (distance matrix ~ n^2/2 bytes): one launch per class keeps the occupancy of
            # the small graphs independent of the largest graph in the batch
            foot = (nn * ((nn + 7) // 8) * 4 + 4 * eo_n + 8 * nn).numpy()
            bounds = [8 << 10, 24 << 10, 56 << 10, 1 << 62]
            classes, lo = [], 0
            for hi in bounds:
                ids = np.nonzero((foot > lo) & (foot <= hi))[0]
                lo = hi
                if ids.size:
                    classes.append((torch.as_tensor(ids.astype(np.int32)).to(src.device), int(nn[ids].max()), int(eo_n[ids].max())))
    d_eptr = edge_ptr.to(dev, non_blocking=True)
    d_nptr = node_ptr.to(dev, non_blocking=True)
    E_in = src.numel()
    counters = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=dev)
    span = _Timer(timings)
    if self_loop:
        if n_total is None:
            n_total = int(node_ptr[-1])
        e_cap = E_in + n_total
        eo = torch.empty((2, e_cap), dtype=torch.int64, device=dev)
        eo_ptr = torch.empty(G + 1, dtype=torch.int64, device=dev)
        tmp = torch.empty(4 * G + 8 * (G // 1024 + 2) + 64, dtype=torch.uint8, device=dev)
        with span('rewrite'):
            _lib.check(L.escgnn_rewrite_self_loops(_ptr(src), _ptr(dst), _ptr(d_eptr), _ptr(d_nptr), G,
                                                   _ptr(eo_ptr), _ptr(eo[0]), _ptr(eo[1]), _ptr(tmp), stream),
                       'rewrite_self_loops')
        eo_src, eo_dst = eo[0], eo[1]
    else:
        e_cap = E_in
        eo_src, eo_dst, eo_ptr = src.contiguous(), dst.contiguous(), d_eptr
        eo = None
    sb = L.escgnn_encode_scratch_bytes(max_nodes, max_edges, int(h))
    if use_rd:
        sb = max(sb, L.escgnn_encode_rd_scratch_bytes(max_nodes, max_edges, int(h)))
    scratch = torch.empty(max(sb, 16), dtype=torch.uint8, device=dev)
    rdh = None
    if use_rd:
        rdh = torch.empty((e_cap + 1, _lib.RD_SLOTS), dtype=torch.int16, device=dev)
        with span('ego_rd'):
            _lib.check(L.escgnn_encode_rd(_ptr(eo_src), _ptr(eo_dst), _ptr(eo_ptr), _ptr(d_nptr), G, int(h),
                                          _ptr(rdh), _ptr(counters), max_nodes, max_edges, _ptr(scratch),
                                          scratch.numel(), stream), 'encode_rd')
    rec_off = torch.empty(e_cap + 1, dtype=torch.int64, device=dev)
    rec_nnz = torch.empty(e_cap + 1, dtype=torch.int32, device=dev)
    edge_graph = torch.empty(e_cap + 1, dtype=torch.int32, device=dev)
    rec_cap = e_cap * 48 + 1024
    for attempt in range(2):
        rec = torch.empty(rec_cap, dtype=torch.int32, device=dev)
        with span('ego_encode'):
            for ids, cn, ce in (classes or [(None, max_nodes, max_edges)]):
                _lib.check(L.escgnn_encode_subset(_ptr(eo_src), _ptr(eo_dst), _ptr(eo_ptr), _ptr(d_nptr),
                                                  ids.numel() if ids is not None else G, _ptr(ids) if ids is not None else None,
                                                  int(h), _ptr(rdh) if rdh is not None else None, _ptr(rec), rec_cap,
                                                  _ptr(rec_off), _ptr(rec_nnz), _ptr(edge_graph), _ptr(counters), cn, ce,
                                                  _ptr(scratch), scratch.numel(), stream), 'encode_subset')
        tail = torch.cat([counters[:2], eo_ptr[G:G + 1]]).cpu()     # the one sync: nnz, error bits, E_out
        nnz, bits, E = int(tail[0]), int(tail[1]), int(tail[2])
        if nnz <= rec_cap:
            break
        rec_cap = nnz
        counters[0] = 0
        counters[2] = 0
    _lib.raise_data_errors(bits)
    res = EncodedBatch(edge_index=(eo[:, :E] if eo is not None else torch.stack([eo_src, eo_dst])), edge_ptr=eo_ptr,
                       rec=rec, rec_off=rec_off[:E], rec_nnz=rec_nnz[:E], edge_graph=edge_graph[:E], num_edges=E,
                       nnz=nnz, use_rd=bool(use_rd))
    if expand:
        out_off = torch.empty(E + 2, dtype=torch.int64, device=dev)
        scan_tmp = torch.empty(E // 1024 + 4, dtype=torch.int64, device=dev)
        with span('scan'):
            _lib.check(L.escgnn_exclusive_scan_i32(_ptr(rec_nnz), E, _ptr(out_off), _ptr(scan_tmp), stream), 'scan')
        trip = torch.empty((3, max(nnz, 1)), dtype=torch.int64, device=dev)
        with span('expand'):
            _lib.check(L.escgnn_expand_records(_ptr(rec), _ptr(rec_off), _ptr(rec_nnz), _ptr(edge_graph),
                                               _ptr(eo_ptr), _ptr(out_off), E, int(use_rd), int(local_ordinals),
                                               _ptr(trip[0]), _ptr(trip[1]), _ptr(trip[2]), stream),
                       'expand_records')
        res.pos_enc, res.pos_index, res.pos_batch = trip[0, :nnz], trip[1, :nnz], trip[2, :nnz]
        res.out_off = out_off[:E + 1]
    return res


def create_subgraphs(data, h=1, sample_ratio=1.0, max_nodes_per_hop=None, node_label='hop', use_rd=False,
                     subgraph_pretransform=None, data_name=None, self_loop=False):
    """Same contract as the reference `create_subgraphs` (utils_edge_efficient.py:20-152).

    `node_label` is accepted and ignored exactly like the reference (both BFS calls hard-code 'hop', :44-51);
    `sample_ratio` / `data_name` are unused there too.  `max_nodes_per_hop` (random sampling, :235-237) and
    `subgraph_pretransform` (k-GNN hook, :109-118) are outside the hot path and must stay None.
    The resistance-distance block follows parity policy E5 (float64)."""
    if max_nodes_per_hop is not None or subgraph_pretransform is not None:
        raise NotImplementedError('max_nodes_per_hop / subgraph_pretransform are not part of the efficient hot path')
    if type(h) == int:
        h = [h]
    assert hasattr(data, 'edge_index') and hasattr(data, 'num_nodes'), 'expected a PyG-style Data object'
    x, edge_index, num_nodes = data.x, data.edge_index, data.num_nodes
    if type(num_nodes) is torch.Tensor:
        num_nodes = num_nodes.item()
    num_nodes = int(num_nodes)
    for h_ in h:
        if h_ >= 5:      # F.one_hot(code, 1300) raises for distances up to h+1 = 6
            raise RuntimeError('Class values must be smaller than num_classes.')
    h_ = int(h[-1])      # the reference returns only the last h (:152 sits outside the loop)
    edge_attr = data.edge_attr
    if self_loop:        # E1 on the attributes (edge_index itself is rewritten on the device)
        keep = edge_index[0] != edge_index[1]
        if edge_attr is not None:
            edge_attr = edge_attr[keep]
            loop_attr = edge_attr.new_full((num_nodes, ) + tuple(edge_attr.size()[1:]), 1.)
            edge_attr = torch.cat([edge_attr, loop_attr], dim=0)
        n_out = int(keep.sum()) + num_nodes
    else:
        n_out = edge_index.size(1)
    if n_out == 0:       # reference: torch.cat of an empty list (:146-151)
        raise RuntimeError('torch.cat(): expected a non-empty list of Tensors')
    if h_ < 1:
        raise NotImplementedError('h must be >= 1')
    ei = edge_index.to(torch.int64)
    if ei.is_cuda:
        eptr = torch.tensor([0, ei.size(1)], dtype=torch.int64)
        nptr = torch.tensor([0, num_nodes], dtype=torch.int64)
        r = encode_batch(ei[0].contiguous(), ei[1].contiguous(), eptr, nptr, h_, use_rd, self_loop,
                         expand=True, local_ordinals=True)
    else:
        r = encode_batch_host(ei[0], ei[1], np.array([0, ei.size(1)], dtype=np.int64),
                              np.array([0, num_nodes], dtype=np.int64), h_, use_rd, self_loop, local_ordinals=True,
                              device=torch.cuda.current_device() if torch.cuda.is_available() else 0)
    kw = dict(pos_enc=r.pos_enc, pos_index=r.pos_index, pos_batch=r.pos_batch)
    if not hasattr(data, 'pos'):
        return data.__class__(data.x, r.edge_index, edge_attr, data.y, None, **kw)
    if not hasattr(data, 'name'):
        return data.__class__(data.x, r.edge_index, edge_attr, data.y, None, **kw)
    return data.__class__(data.x, r.edge_index, edge_attr, data.y, pos=data.pos, name=data.name,
                          node_type=data.node_type, **kw)


# ------------------------------------------------------------------------------------------------------------------
# GraphGPS twin of the transform: the same encodings plus `attn_bias`, the flattened
# all-pairs shortest-path matrix the transformer layers use as an attention bias.
def all_pairs_spd_batch(src, dst, edge_ptr, node_ptr, unreachable=100, device=None):
    """Shortest-path lengths between all node pairs of every graph (undirected, unreachable -> `unreachable`).
    src/dst: int64 graph-local node ids, edge_ptr/node_ptr: int64 [G+1] (host or device).  Returns
    (flat int64 tensor on the device, out_ptr int64 [G+1] on the host) -- graph g's [n, n] block is
    flat[out_ptr[g]:out_ptr[g+1]].view(n, n).  Reference: GraphGPS/graphgps/loader/utils_escgnn.py:29-38."""
    if not torch.cuda.is_available():
        raise RuntimeError('esc_gnn_b200: no CUDA device -- the transform has no CPU fallback')
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    ep = torch.as_tensor(edge_ptr, dtype=torch.int64).cpu()
    npt = torch.as_tensor(node_ptr, dtype=torch.int64).cpu()
    n = npt[1:] - npt[:-1]
    out_ptr = torch.zeros(npt.numel(), dtype=torch.int64)
    out_ptr[1:] = torch.cumsum(n * n, 0)
    G = int(n.numel())
    max_n = int(n.max()) if G else 0
    max_e = int((ep[1:] - ep[:-1]).max()) if G else 0
    L = _lib.lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    s, d = torch.as_tensor(src).to(dev, torch.int64).contiguous(), torch.as_tensor(dst).to(dev, torch.int64).contiguous()
    out = torch.empty(int(out_ptr[-1]), dtype=torch.int64, device=dev)
    counters = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=dev)
    if G:
        ep_d, np_d, op_d = ep.to(dev), npt.to(dev), out_ptr.to(dev)       # named: they must outlive the launch
        _lib.check(L.escgnn_all_pairs_spd(P(s), P(d), P(ep_d), P(np_d), G, P(op_d), P(out), max_n, max_e,
                                          int(unreachable), P(counters),
                                          ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), 'all_pairs_spd')
        _lib.raise_data_errors(int(counters[1]))
    return out, out_ptr


def all_pairs_spd(edge_index, num_nodes, unreachable=100):
    """`attn_bias` of ONE graph: int64 [num_nodes * num_nodes], on the device of `edge_index`."""
    ei = edge_index.to(torch.int64)
    out, _ = all_pairs_spd_batch(ei[0], ei[1], [0, ei.size(1)], [0, int(num_nodes)], unreachable,
                                 device=ei.device if ei.is_cuda else None)
    return out.to(edge_index.device)


def create_subgraphs_gps(data, h=1, sample_ratio=1.0, max_nodes_per_hop=None, node_label='hop', use_rd=False,
                         subgraph_pretransform=None, data_name=None, self_loop=False):
    """`create_subgraphs` of the GraphGPS loader (GraphGPS/graphgps/loader/utils_escgnn.py:22-154): the encodings of
    `create_subgraphs` plus `attn_bias` computed on the INPUT graph (before self-loops are appended, :29-38)."""
    num_nodes = data.num_nodes
    num_nodes = int(num_nodes.item()) if torch.is_tensor(num_nodes) else int(num_nodes)
    out = create_subgraphs(data, h, sample_ratio, max_nodes_per_hop, node_label, use_rd, subgraph_pretransform, data_name,
                           self_loop)
    out.attn_bias = all_pairs_spd(data.edge_index, num_nodes)
    return out
