"""Autograd bindings of the sm_100a model kernels (C-ABI in include/escgnn_b200.h), on the current torch stream.

Every op raises when its inputs are not CUDA tensors: there is no CPU / eager fallback behind these functions.
"""
import ctypes

import torch

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


TIMINGS = None      # bench.py sets this to a dict to bracket every kernel call with CUDA events


class _span(object):
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if TIMINGS is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if TIMINGS is not None:
            self.b.record()
            TIMINGS.setdefault(self.name, []).append((self.a, self.b))


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('esc_gnn_b200 ops need CUDA tensors; there is no CPU fallback')


class GraphIndex(object):
    """Per-batch index structures built once on the device and shared by every layer:
    CSR by destination (forward aggregation), CSR by source (backward), record ranges per edge, node ranges per graph."""

    def __init__(self, edge_index, num_nodes, batch=None, num_graphs=None, pos_batch=None):
        _need_cuda(edge_index)
        L = _lib.lib()
        dev = edge_index.device
        st = _stream(edge_index)
        ei = edge_index.contiguous()
        self.src, self.dst = ei[0], ei[1]
        self.num_nodes, self.num_edges = int(num_nodes), int(ei.size(1))
        N, E = self.num_nodes, self.num_edges
        buf = torch.empty(4 * (N + 1) + 2 * E + 2, dtype=torch.int32, device=dev)
        self.dst_ptr, self.src_ptr = buf[:N + 1], buf[N + 1:2 * N + 2]
        tmp_a, tmp_b = buf[2 * N + 2:3 * N + 3], buf[3 * N + 3:4 * N + 4]
        self.dst_perm, self.src_perm = buf[4 * N + 4:4 * N + 4 + E], buf[4 * N + 4 + E:4 * N + 4 + 2 * E]
        self.err = torch.zeros(1, dtype=torch.int64, device=dev)
        _lib.check(L.escgnn_csr_build(_p(self.dst), E, N, _p(self.dst_ptr), _p(self.dst_perm), _p(tmp_a),
                                      _p(self.err), None, st), 'csr_build')
        _lib.check(L.escgnn_csr_build(_p(self.src), E, N, _p(self.src_ptr), _p(self.src_perm), _p(tmp_b),
                                      _p(self.err), None, st), 'csr_build')
        self.rec_ptr = None
        if pos_batch is not None:
            self.rec_ptr = torch.empty(E + 1, dtype=torch.int32, device=dev)
            _lib.check(L.escgnn_sorted_ids_to_ptr(_p(pos_batch), pos_batch.numel(), E, _p(self.rec_ptr), None, st),
                       'sorted_ids_to_ptr')
        self.graph_ptr, self.num_graphs = None, None
        if batch is not None:
            if num_graphs is None:
                num_graphs = int(batch[-1]) + 1          # the reference's own sync (PyG global_add_pool)
            self.num_graphs = int(num_graphs)
            self.graph_ptr = torch.empty(self.num_graphs + 1, dtype=torch.int32, device=dev)
            _lib.check(L.escgnn_sorted_ids_to_ptr(_p(batch), batch.numel(), self.num_graphs, _p(self.graph_ptr), None,
                                                  st),
                       'sorted_ids_to_ptr')


def _field(data, name):
    """Optional attribute of a Data-like object (our Data keeps fields in `_store`, PyG's raises AttributeError)."""
    store = getattr(data, '_store', None)
    if isinstance(store, dict):
        return store.get(name)
    try:
        return getattr(data, name)
    except AttributeError:
        return None


def graph_index(data):
    """Index of a batch, cached on the batch object (edge_index / pos_batch / batch are immutable per batch)."""
    idx = data.__dict__.get('_esc_index') if hasattr(data, '__dict__') else None
    if idx is not None and idx.src.data_ptr() == data.edge_index[0].data_ptr():
        return idx
    n = data.x.size(0)
    try:
        ng = data.num_graphs
    except AttributeError:
        ng = None
    idx = GraphIndex(data.edge_index, n, batch=data.batch, num_graphs=ng,
                     pos_batch=None if _field(data, 'rec') is not None else data.pos_batch)
    try:
        object.__setattr__(data, '_esc_index', idx)
    except Exception:
        pass
    return idx


class _BagEmbed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, pos_index, pos_enc, rec_ptr, n_edges):
        _need_cuda(weight, pos_index, pos_enc, rec_ptr)
        H = weight.size(1)
        out = torch.empty((n_edges, H), dtype=torch.float32, device=weight.device)
        w = weight.contiguous()
        with _span('bag_embed_fwd'):
            _lib.check(_lib.lib().escgnn_bag_embed_fwd(_p(w), H, _p(pos_index), _p(pos_enc), _p(rec_ptr), None, None, None,
                                                       n_edges, _p(out), None, _stream(w)), 'bag_embed_fwd')
        ctx.save_for_backward(pos_index, pos_enc, rec_ptr)
        ctx.shape = tuple(weight.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        pos_index, pos_enc, rec_ptr = ctx.saved_tensors
        g = g.contiguous()
        dW = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        with _span('bag_embed_bwd'):
            _lib.check(_lib.lib().escgnn_bag_embed_bwd(_p(g), ctx.shape[1], _p(pos_index), _p(pos_enc), _p(rec_ptr), None,
                                                       None, None, g.size(0), _p(dW), None, _stream(g)), 'bag_embed_bwd')
        return dW, None, None, None, None


class _BagEmbedRec(torch.autograd.Function):
    """Bag-embed straight from the encoder's packed records (no int64 triple in between)."""
    @staticmethod
    def forward(ctx, weight, rec, rec_off, rec_nnz, n_edges):
        _need_cuda(weight, rec, rec_off, rec_nnz)
        H = weight.size(1)
        out = torch.empty((n_edges, H), dtype=torch.float32, device=weight.device)
        w = weight.contiguous()
        with _span('bag_embed_fwd'):
            _lib.check(_lib.lib().escgnn_bag_embed_fwd(_p(w), H, None, None, None, _p(rec), _p(rec_off), _p(rec_nnz),
                                                       n_edges, _p(out), None, _stream(w)), 'bag_embed_fwd')
        ctx.save_for_backward(rec, rec_off, rec_nnz)
        ctx.shape = tuple(weight.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        rec, rec_off, rec_nnz = ctx.saved_tensors
        g = g.contiguous()
        dW = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        with _span('bag_embed_bwd'):
            _lib.check(_lib.lib().escgnn_bag_embed_bwd(_p(g), ctx.shape[1], None, None, None, _p(rec), _p(rec_off),
                                                       _p(rec_nnz), g.size(0), _p(dW), None, _stream(g)), 'bag_embed_bwd')
        return dW, None, None, None, None


def bag_embed_data(weight, data, index):
    """Dispatch on what the batch carries: packed encoder records (native path) or the reference's int64 triple."""
    rec = _field(data, 'rec')
    if rec is not None:
        return _BagEmbedRec.apply(weight, rec, data.rec_off, data.rec_nnz, index.num_edges)
    return bag_embed(weight, data.pos_index, data.pos_enc, index)


def bag_embed(weight, pos_index, pos_enc, index):
    """z0[e] = sum_k pos_enc[k] * weight[pos_index[k]] over the records of edge e (run_graphcount.py:155)."""
    return _BagEmbed.apply(weight, pos_index.contiguous(), pos_enc.contiguous(), index.rec_ptr, index.num_edges)


class _GineAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, edge_feat, eps, index):
        _need_cuda(x, edge_feat, eps)
        x, edge_feat = x.contiguous(), edge_feat.contiguous()
        N, C = x.shape
        assert edge_feat.shape == (index.num_edges, C) and N == index.num_nodes
        out = torch.empty_like(x)
        with _span('gine_aggregate_fwd'):
            _lib.check(_lib.lib().escgnn_gine_aggregate_fwd(_p(x), _p(edge_feat), _p(index.src), _p(index.dst_ptr),
                                                            _p(index.dst_perm), _p(eps), N, C, _p(out), None, _stream(x)),
                       'gine_aggregate_fwd')
        ctx.save_for_backward(x, edge_feat, eps)
        ctx.index = index
        return out

    @staticmethod
    def backward(ctx, g):
        x, edge_feat, eps = ctx.saved_tensors
        index = ctx.index
        g = g.contiguous()
        N, C = x.shape
        gx = torch.empty_like(x)
        ge = torch.empty_like(edge_feat)
        dots = torch.empty(N, dtype=torch.float32, device=x.device)
        geps = torch.empty(1, dtype=torch.float32, device=x.device)
        with _span('gine_aggregate_bwd'):
            _lib.check(_lib.lib().escgnn_gine_aggregate_bwd(_p(g), _p(x), _p(edge_feat), _p(index.dst), _p(index.src_ptr),
                                                            _p(index.src_perm), _p(eps), N, C, _p(gx), _p(ge), _p(dots),
                                                            _p(geps), None, _stream(x)), 'gine_aggregate_bwd')
        return gx, ge, geps, None


def gine_aggregate(x, edge_feat, eps, index):
    """(1+eps) x[i] + sum_{e: dst=i} relu(x[src_e] + edge_feat[e])  -- PyG GINEConv propagate + residual."""
    return _GineAggregate.apply(x, edge_feat, eps, index)


class _SegmentPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ptr, segs, mean):
        _need_cuda(x, ptr)
        x = x.contiguous()
        out = torch.empty((segs, x.size(1)), dtype=torch.float32, device=x.device)
        with _span('segment_pool_fwd'):
            _lib.check(_lib.lib().escgnn_segment_pool_fwd(_p(x), _p(ptr), segs, x.size(1), int(mean), _p(out), _stream(x)),
                       'segment_pool_fwd')
        ctx.save_for_backward(ptr)
        ctx.meta = (x.size(0), x.size(1), segs, int(mean))
        return out

    @staticmethod
    def backward(ctx, g):
        (ptr, ) = ctx.saved_tensors
        n, c, segs, mean = ctx.meta
        g = g.contiguous()
        gx = torch.zeros((n, c), dtype=torch.float32, device=g.device)
        with _span('segment_pool_bwd'):
            _lib.check(_lib.lib().escgnn_segment_pool_bwd(_p(g), _p(ptr), segs, c, mean, _p(gx), _stream(g)),
                       'segment_pool_bwd')
        return gx, None, None, None


def global_add_pool(x, index):
    return _SegmentPool.apply(x, index.graph_ptr, index.num_graphs, False)


def global_mean_pool(x, index):
    return _SegmentPool.apply(x, index.graph_ptr, index.num_graphs, True)


# ---------------------------------------------------------------------------------------------------------------------
# Dense layers of the drop-in modules on the hand-written kernels (same parameters / state_dict keys as torch.nn)
_GEMM_WS = {}


def _gemm_ws(dev):
    ws = _GEMM_WS.get(dev)
    if ws is None:
        ws = _GEMM_WS[dev] = torch.empty(4 * 1024 * 1024, dtype=torch.float32, device=dev)
    return ws


def _gemm(A, a_mn, B, b_mn, C, bias, M, N, K, accumulate=False):
    """C[M,N] (+)= op(A) op(B)^T (+bias): tcgen05 3xTF32 kernel when TMA alignment allows, CUDA-core kernel otherwise."""
    L = _lib.lib()
    st = _stream(C)
    if all(t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0 for t in (A, B)):
        ws = _gemm_ws(C.device)
        _lib.check(L.escgnn_gemm_tf32x3(_p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), _p(C), C.stride(0), _p(bias),
                                        M, N, K, int(accumulate), _p(ws), ws.numel(), st), 'gemm_tf32x3')
    else:
        _lib.check(L.escgnn_gemm_simple(_p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), _p(C), C.stride(0), _p(bias),
                                        M, N, K, int(accumulate), st), 'gemm_simple')


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        _need_cuda(x, weight)
        x = x.contiguous()
        M, K = x.shape
        N = weight.size(0)
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        if M:
            _gemm(x, False, weight, False, y, bias, M, N, K)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        M, K = x.shape
        N = weight.size(0)
        dx = torch.empty_like(x) if M else torch.zeros_like(x)
        dw = torch.empty_like(weight) if M else torch.zeros_like(weight)
        if M:
            _gemm(dy, False, weight, True, dx, None, M, K, N)        # dX = dY W        (W read as an MN-major operand)
            _gemm(dy, True, x, True, dw, None, N, K, M)              # dW = dY^T X      (both operands MN-major, split-K)
        db = dy.sum(0) if ctx.has_bias else None
        return dx, dw, db


class Linear(torch.nn.Linear):
    """torch.nn.Linear whose forward / dgrad / wgrad run on the tcgen05 3xTF32 GEMM (csrc/gemm_tf32x3.cu)."""
    def forward(self, x):
        if x.dim() != 2 or not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError('esc_gnn_b200.ops.Linear needs a 2-D fp32 CUDA input; there is no CPU fallback')
        return _LinearFn.apply(x, self.weight, self.bias)


_ROWS, _PARTIAL = {}, {}


def _device_rows(rows, dev):
    """The row count as a 1-element device tensor, cached: building it per call is a pageable host-to-device copy that
    stalls the launching thread."""
    key = (dev, int(rows))
    t = _ROWS.get(key)
    if t is None:
        if len(_ROWS) > 4096:
            _ROWS.clear()
        t = _ROWS[key] = torch.tensor([rows], dtype=torch.int32, device=dev)
    return t


def _reduction_workspace(rows, C, dev):
    """One zero-initialised reduction workspace per device, grown on demand (the kernels leave its ticket words at zero and
    calls on one stream are serialised)."""
    need = _lib.lib().escgnn_dense_partial_floats(rows, C)
    ws = _PARTIAL.get(dev)
    if ws is None or ws.numel() < need:
        ws = _PARTIAL[dev] = torch.zeros(max(need, 1 << 20), dtype=torch.float32, device=dev)
    return ws


class _BatchNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, training):
        _need_cuda(x, weight)
        x = x.contiguous()
        rows, C = x.shape
        L = _lib.lib()
        d_rows = _device_rows(rows, x.device)
        partial = _reduction_workspace(rows, C, x.device)
        mean, rstd = torch.empty(C, device=x.device), torch.empty(C, device=x.device)
        y = torch.empty_like(x)
        _lib.check(L.escgnn_bn_act_fwd(_p(x), C, _p(weight), _p(bias), _p(running_mean), _p(running_var), _p(mean), _p(rstd),
                                       _p(partial), 0, eps, momentum, int(training), _p(d_rows), rows, C, _p(y), C, _stream(x)),
                   'bn_act_fwd')
        ctx.save_for_backward(x, weight, bias, mean, rstd, d_rows, partial)
        ctx.training = training
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, mean, rstd, d_rows, partial = ctx.saved_tensors
        dy = dy.contiguous()
        rows, C = x.shape
        dx, dg, db = torch.empty_like(x), torch.empty(C, device=x.device), torch.empty(C, device=x.device)
        _lib.check(_lib.lib().escgnn_bn_act_bwd(_p(x), C, _p(dy), C, None, 0, _p(mean), _p(rstd), _p(weight), _p(bias), 0,
                                                int(ctx.training), _p(partial), _p(d_rows), rows, C, _p(dg), _p(db), _p(dx), C,
                                                _stream(x)), 'bn_act_bwd')
        return dx, dg, db, None, None, None, None, None


class BatchNorm1d(torch.nn.BatchNorm1d):
    """torch.nn.BatchNorm1d (affine, running stats, fixed momentum) on the bn_act kernels of csrc/dense_ops.cu."""
    def forward(self, x):
        if x.dim() != 2 or not x.is_cuda:
            raise RuntimeError('esc_gnn_b200.ops.BatchNorm1d needs a 2-D CUDA input; there is no CPU fallback')
        if self.training:
            self.num_batches_tracked.add_(1)
        if x.size(0) == 0:
            return x
        return _BatchNormFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps,
                                  self.momentum if self.momentum is not None else 0.1, self.training)
