"""Static-shape training engine: the whole hot path of one step as ONE replayable CUDA graph.

    raw graphs (static input buffers) -> E1/E5/E2-E4 encoding -> device collation -> CSR builds ->
    NestedGIN_eff forward -> loss -> hand-written backward -> (gradient all-reduce) -> Adam

Why: at the reference's batch sizes (256 graphs = ~6k nodes, ~12k edges) a step is a few hundred microseconds of
GPU work, and an eager autograd step spends >85% of its time in launch latency.  Every buffer here is
allocated once at a fixed capacity, every kernel reads the actual sizes (nodes, edges, graphs, records) from a
4-int device array, so the captured graph is valid for every batch and the host does nothing per step but copy the
raw graphs into the input buffers and replay.

The engine reads and updates the SAME parameters as the drop-in `NestedGIN_eff` module it is built from (they are
views into FlatAdam's flat buffer), so `state_dict()` / checkpoints stay interchangeable with the reference's
(run_zinc.py:257-262, run_graphcount.py:464-476).  Forward semantics: zinc_models.py:579-611 and
run_graphcount.py:134-194; train step: run_zinc.py:266-289.

Dense contractions (every nn.Linear forward / dgrad / wgrad) run on the hand-written tcgen05 3xTF32 GEMM
(csrc/gemm_tf32x3.cu: TMA -> shared memory -> tcgen05.mma -> TMEM); everything else is the hand-written sm_100a
kernels of csrc/*.cu.  No library GEMM is on the path.
"""
import ctypes
import os

import torch

from . import _lib
from .optim import FlatAdam

ACT = {'none': 0, 'relu': 1, 'elu': 2}


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _Ctx(object):
    """Allocation + launch helpers shared by the tapes."""

    def __init__(self, device, caps, dims):
        self.dev = device
        self.caps = dict(caps)                       # 'N', 'E', 'B', 'nnz', 'E_in'
        self.dims = dims                             # int32[4] on the device: N, E, B, nnz of the batch being trained on
        self.rows = {'N': self.dims[0:1], 'E': self.dims[1:2], 'B': self.dims[2:3]}
        self.partial = torch.zeros(_lib.lib().escgnn_dense_partial_floats(max(self.caps['N'], self.caps['E'], self.caps['B']), 2048),
                                   dtype=torch.float32, device=device)
        self.L = _lib.lib()

    def st(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def buf(self, kind, cols, dtype=torch.float32):
        return torch.zeros((self.caps[kind], cols), dtype=dtype, device=self.dev)


class _BatchSet(object):
    """Everything the training tapes read about ONE encoded batch, carved out of a single contiguous allocation so that a
    whole batch moves from the encoder's staging set to the live set with one device-to-device copy (pipelined mode)."""

    def __init__(self, dev, variant, G, N, E, nnz_cap):
        i64, i32, f32 = torch.int64, torch.int32, torch.float32
        spec = [('rec', i32, (nnz_cap, )), ('rec_off', i64, (E + 1, )), ('rec_nnz', i32, (E + 1, )), ('ei', i64, (2, E)),
                ('batch', i64, (N, )), ('idx_buf', i32, (4 * (N + 1) + 2 * E + 2, )), ('graph_ptr', i32, (G + 1, )),
                ('dims', i32, (4, ))]
        if variant == 'zinc':
            spec += [('in_x', i64, (N, )), ('in_ea', i64, (E, )), ('in_y', f32, (G, ))]
        elif variant == 'ogb':          # 9 atom columns, 3 bond columns (after E1), one target per graph (NaN = unlabelled)
            spec += [('in_x', i64, (N, 9)), ('in_ea', i64, (E, 3)), ('in_y', f32, (G, ))]
        else:
            spec += [('in_x', f32, (N, 10)), ('in_y', f32, (N, ))]
        esz = {i64: 8, i32: 4, f32: 4}
        off, offs = 0, []
        for _, dt, shape in spec:
            offs.append(off)
            n = 1
            for d in shape:
                n *= d
            off += (n * esz[dt] + 255) // 256 * 256
        self.buf = torch.zeros(off, dtype=torch.uint8, device=dev)
        for (name, dt, shape), o in zip(spec, offs):
            n = 1
            for d in shape:
                n *= d
            setattr(self, name, self.buf[o:o + n * esz[dt]].view(dt).view(shape))
        if variant == 'count':
            self.in_ea = None
        b = self.idx_buf
        self.dst_ptr, self.src_ptr = b[:N + 1], b[N + 1:2 * N + 2]
        self.tmp_a, self.tmp_b = b[2 * N + 2:3 * N + 3], b[3 * N + 3:4 * N + 4]
        self.dst_perm, self.src_perm = b[4 * N + 4:4 * N + 4 + E], b[4 * N + 4 + E:4 * N + 4 + 2 * E]
        self.rows = {'N': self.dims[0:1], 'E': self.dims[1:2], 'B': self.dims[2:3]}


class StaticTrainEngine(object):
    """One model variant at a fixed capacity: NestedGIN_eff 'zinc' / 'count', or 'ogb' = GNN(gnn_type='gin_eff')."""

    def __init__(self, model, variant, flags, max_graphs, max_nodes_per_graph, max_edges_per_graph, nodes_cap, edges_cap,
                 lr=1e-3, distributed=False, records_per_edge=64, use_graph=True, tensor_cores=True, pipeline=False, atomic_wgrad=True, encoder_ctas=None, fused_head=True, fuse_bn=False,
                 exchange='nccl'):
        if variant not in ('zinc', 'count', 'ogb'):
            raise NotImplementedError('engine variants: zinc, count, ogb')
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise RuntimeError('StaticTrainEngine needs a CUDA model; there is no CPU fallback')
        self.model, self.variant, self.flags = model, variant, dict(flags)
        # data-parallel exchange: 'p2p' = reduce-scatter + Adam + all-gather as one kernel over NVLink peer memory, a node of the
        # step's single graph (csrc/p2p.cu); 'nccl' = one all-reduce of the flat gradient between two captured graphs
        if exchange not in ('nccl', 'p2p', 'auto'):
            raise ValueError("exchange: 'nccl', 'p2p' or 'auto' (p2p when the peers' memory can be mapped, else nccl)")
        self.exchange = exchange if distributed else 'none'
        p2p_group = None if self.exchange in ('p2p', 'auto') else False
        # the edge projections `conv.lin` of ALL layers read the same z, so they are one GEMM against the row-concatenation
        # of their weights: lay those tensors out adjacently (256-wide layers first, the narrow first layer last)
        if variant == 'ogb':
            gn = model.gnn_node
            if gn.JK != 'last' or model.num_tasks != 1 or not atomic_wgrad:
                raise NotImplementedError('ogb engine variant: JK="last", one task, atomic_wgrad=True')
            # adjacent in the flat buffer: the per-layer projections of z (one grouped GEMM), the 9 atom tables (one lookup with
            # row offsets), and every layer's 3 bond tables
            first = [cv.edge_encoder_pos.weight for cv in gn.convs] + [cv.edge_encoder_pos.bias for cv in gn.convs]
            first += [e.weight for e in gn.node_encoder.atom_embedding_list]
            for cv in gn.convs:
                first += [e.weight for e in cv.edge_encoder.bond_embedding_list]
            last = []
        else:
            self.lin_convs = list(model.convs) + [model.conv1]
            first = [cv.lin.weight for cv in self.lin_convs] + [cv.lin.bias for cv in self.lin_convs]
            # parameters whose gradients are produced by the LAST kernels of the backward pass (the z path, the input embeddings):
            # one contiguous tail of the flat vector = the late bucket of the peer-memory exchange; everything before it is
            # exchanged on the side branch while those kernels still run
            last = [model.z_initial.weight] + list(model.z_embedding.parameters())
            last += list(model.x_embedding.parameters()) if variant == 'count' else \
                [model.node_type_embedding.weight, model.edge_type_embedding.weight]
        try:
            self.opt = FlatAdam(model.parameters(), lr=lr, first=first, p2p_group=p2p_group, last=last)
        except RuntimeError:
            if self.exchange != 'auto':
                raise
            self.opt = FlatAdam(model.parameters(), lr=lr, first=first, p2p_group=False, last=last)    # (collective failure: every rank lands here)
        if self.exchange == 'auto':
            self.exchange = 'p2p' if self.opt.peers is not None else 'nccl'
        self.distributed, self.use_graph = distributed, use_graph
        dev = p0.device
        self.G = int(max_graphs)
        self.max_n, self.max_e = int(max_nodes_per_graph), int(max_edges_per_graph)   # per-graph maxima (after E1)
        e_in_cap = int(edges_cap)
        e_cap = e_in_cap + (int(nodes_cap) if flags['self_loop'] else 0)
        caps = dict(N=int(nodes_cap), E=e_cap, B=self.G, nnz=e_cap * records_per_edge, E_in=e_in_cap)
        # pipeline: step(raw_k) trains on batch k-1 while the encoder works on batch k (two branches of one graph); the encoder
        # then writes a staging set that becomes live with one copy at the start of the next step
        self.pipeline = bool(pipeline)
        self.live = _BatchSet(dev, variant, self.G, caps['N'], caps['E'], caps['nnz'])
        self.stage = _BatchSet(dev, variant, self.G, caps['N'], caps['E'], caps['nnz']) if self.pipeline else self.live
        self.c = c = _Ctx(dev, caps, self.live.dims)
        i64 = torch.int64
        # ---- static raw-input buffers (the only thing the host touches per step)
        self.in_src = torch.zeros(e_in_cap, dtype=i64, device=dev)
        self.in_dst = torch.zeros(e_in_cap, dtype=i64, device=dev)
        self.in_eptr = torch.zeros(self.G + 1, dtype=i64, device=dev)
        self.in_nptr = torch.zeros(self.G + 1, dtype=i64, device=dev)
        # ---- encoder-private state
        self.counters = torch.zeros(_lib.NUM_COUNTERS, dtype=i64, device=dev)
        if flags['self_loop']:
            self.eo = torch.zeros((2, e_cap), dtype=i64, device=dev)
            self.eo_ptr = torch.zeros(self.G + 1, dtype=i64, device=dev)
            self.rw_tmp = torch.zeros(4 * self.G + 8 * (self.G // 1024 + 2) + 64, dtype=torch.uint8, device=dev)
        sb = c.L.escgnn_encode_scratch_bytes(self.max_n, self.max_e, flags['h'])
        if flags['use_rd']:
            sb = max(sb, c.L.escgnn_encode_rd_scratch_bytes(self.max_n, self.max_e, flags['h']))
        self.scratch = torch.zeros(max(sb, 16), dtype=torch.uint8, device=dev)
        self.rdh = torch.zeros((e_cap + 1, _lib.RD_SLOTS), dtype=torch.int16, device=dev) if flags['use_rd'] else None
        self.edge_graph = torch.zeros(e_cap + 1, dtype=torch.int32, device=dev)
        # ---- what the tapes read: the LIVE batch set
        lv = self.live
        self.in_x, self.in_ea, self.in_y = lv.in_x, lv.in_ea, lv.in_y
        self.rec, self.rec_off, self.rec_nnz = lv.rec, lv.rec_off, lv.rec_nnz
        self.ei, self.batch, self.graph_ptr = lv.ei, lv.batch, lv.graph_ptr
        self.dst_ptr, self.src_ptr, self.dst_perm, self.src_perm = lv.dst_ptr, lv.src_ptr, lv.dst_perm, lv.src_perm
        self.enc_stream = torch.cuda.Stream(device=dev)
        self._primed = False
        if self.pipeline:
            class _Raw(object):
                pass
            self.raw_feat = _Raw()
            self.raw_feat.in_x = torch.zeros_like(lv.in_x)
            self.raw_feat.in_y = torch.zeros_like(lv.in_y)
            self.raw_feat.in_ea = torch.zeros_like(lv.in_ea) if lv.in_ea is not None else None
        self.idx_err = torch.zeros(1, dtype=i64, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.tensor_cores = tensor_cores
        # weight gradients land in the flat gradient buffer, which every step zeroes first: their split-K slices can be added
        # in place with vector reductions (2) instead of going through partial tiles and a reduction launch (0)
        self.wgrad_mode = 2 if atomic_wgrad else 0
        self.bounded_gemm = True
        self.bucketed_exchange, self._early_done, self._xchg_done = True, False, None
        self._loss_host = None                   # step_read(): pinned slots for the host's copy of the loss
        self.xchg = torch.cuda.Stream(device=dev)
        # Linear -> BatchNorm -> activation as one launch each way (GEMM epilogues behind a grid barrier); only kernels of the MAIN
        # branch take that path (two barrier kernels on concurrent branches could starve each other of SM slots)
        self.fuse_bn = bool(fuse_bn) and tensor_cores
        self.bn_ws = torch.zeros(int(self.c.L.escgnn_linear_bn_workspace_floats(max(caps['N'], caps['E'], caps['B']), 2048)),
                                 dtype=torch.float32, device=dev)
        self.fused_head = fused_head
        self.encoder_ctas = int(encoder_ctas) if encoder_ctas else None
        self.gemm_ws = torch.zeros(8 * 1024 * 1024, dtype=torch.float32, device=dev)    # split-K partial tiles (wgrad)
        # weight / bias gradients are off the critical path (only Adam consumes them): they run on a side stream that
        # forks from the backward chain wherever a dY becomes available and joins before the optimiser
        self.gemm_ws_side = torch.zeros_like(self.gemm_ws)
        self.side = torch.cuda.Stream(device=dev)
        self.side_partial = torch.zeros_like(c.partial)
        self._side_used = False
        # the z path's share of the backward pass starts with d z = sum over layers of (d edge projection) W_l.  Opt-in
        # (ESCGNN_SPLIT_PROJ=1): each layer's term is formed on a third branch as soon as that layer's aggregation backward has
        # produced it, so only the first layer's (narrow) term is left when the layer chain ends.  Measured SLOWER at batch 256
        # (1.228 vs 1.159 ms per step): the 190-CTA edge-level products evict the node-level GEMMs of the critical chain from the SMs,
        # which costs more than the 55 us the shorter tail saves.  Default: one product over all layers at the end.
        self.split_proj = os.environ.get('ESCGNN_SPLIT_PROJ', '0') == '1'
        self.proj_stream = torch.cuda.Stream(device=dev)
        self.gemm_ws_proj = torch.zeros(1024 * 1024, dtype=torch.float32, device=dev)
        self.inline_branches = False
        self.fwd, self.bwd, self._bns, self._emb_ready = [], [], [], []
        # bond attributes go through E1 on the device whenever the edge list does (dropped loops, appended loop rows = 1:
        # utils_edge_efficient.py:35-36 / PyG add_self_loops fill value) -- always for 'ogb', for 'zinc' with self_loop
        self.ea_cols = 3 if variant == 'ogb' else 1
        self.ea_via_raw = variant == 'ogb' or (variant == 'zinc' and bool(flags['self_loop']))
        if self.ea_via_raw:
            shape = (e_in_cap, 3) if variant == 'ogb' else (e_in_cap, )
            self.in_ea_raw = torch.zeros(shape, dtype=torch.int64, device=dev)     # bond columns before E1
        if variant == 'ogb':
            self._build_ogb_tape()
        else:
            self._build_model_tape()
        self._bn_synced = 0
        self.graph = None
        self.steps = 0

    # ------------------------------------------------------------------ primitive ops (append to the tapes)
    # ---- dense contractions: tcgen05 3xTF32 GEMM (csrc/gemm_tf32x3.cu); CUDA-core kernel for 10-wide odd shapes
    def _gemm(self, tag, A, a_mn, B, b_mn, C, bias, M, N, K, accumulate, rows=None):
        """rows: the capacity dimension of this product (M for forward / dgrad, K for wgrad) as a row kind 'N' / 'E' / 'B':
        the kernel then follows the batch's actual row count (read on the device) instead of the capacity."""
        c = self.c
        ok = all(t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0 for t in (A, B)) and self.tensor_cores
        if ok:
            # split-K partials: one workspace per stream (GEMMs of the two graph branches may run concurrently)
            cur = torch.cuda.current_stream(c.dev)
            ws = self.gemm_ws_side if cur == self.side else self.gemm_ws_proj if cur == self.proj_stream else self.gemm_ws
            d_rows = _p(c.rows[rows]) if (rows is not None and self.bounded_gemm) else None
            _lib.check(c.L.escgnn_gemm_tf32x3_bounded(_p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), _p(C),
                                                      C.stride(0), _p(bias), M, N, K, int(accumulate), _p(ws), ws.numel(), d_rows,
                                                      (2 if tag == 'gemm_wgrad' else 1) | 4, c.st()), tag)      # | 4: counts final long before
        else:
            _lib.check(c.L.escgnn_gemm_simple(_p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), _p(C), C.stride(0),
                                              _p(bias), M, N, K, int(accumulate), c.st()), tag + '_simple')

    # ---- graph branches: work that is off the critical path runs on the side stream (captured as a parallel branch)
    def _fork(self, fn, stream=None):
        """Run fn on the side stream (or `stream`) after everything issued so far on the main stream; returns its completion event.
        Work forked to the side stream is joined by _join(); the caller waits for the event of any other stream itself."""
        main = torch.cuda.current_stream(self.c.dev)
        if self.inline_branches:                  # profiling: one stream, so every interval belongs to exactly one kernel
            fn()
            done = torch.cuda.Event()
            done.record(main)
            return done
        side = self.side if stream is None else stream
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            fn()
            done = torch.cuda.Event()
            done.record(side)
        if stream is None:
            self._side_used = True
        return done

    def _join(self):
        if self._side_used:
            done = torch.cuda.Event()
            done.record(self.side)
            torch.cuda.current_stream(self.c.dev).wait_event(done)
            self._side_used = False

    def _tc_ok(self, *ts):
        return self.tensor_cores and all(t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0 for t in ts)

    def _bn_spec(self, bn, act, kind, x=None, dx=None, out=None):
        """One training-mode BatchNorm + activation of the tape: x (pre-BN) -> out; dx = gradient wrt x.  mean / rstd are saved by
        whichever kernel runs its forward (stand-alone or a GEMM epilogue) for whichever runs its backward."""
        C = bn.num_features
        return dict(bn=bn, act=act, kind=kind, x=x, dx=dx, out=out,
                    mean=torch.zeros(C, dtype=torch.float32, device=self.c.dev), rstd=torch.ones(C, dtype=torch.float32, device=self.c.dev))

    def _bn_fwd(self, s):
        """Stand-alone forward of a BatchNorm spec (its backward may still be fused into the next Linear's dgrad)."""
        c, bn, x, out, kind = self.c, s['bn'], s['x'], s['out'], s['kind']
        self.fwd.append(lambda: _lib.check(c.L.escgnn_bn_act_fwd(
            _p(x), x.stride(0), _p(bn.weight), _p(bn.bias), _p(bn.running_mean), _p(bn.running_var), _p(s['mean']), _p(s['rstd']),
            _p(c.partial), ACT[s['act']], bn.eps, bn.momentum, 1, _p(c.rows[kind]), c.caps[kind], x.size(1), _p(out), out.stride(0),
            c.st()), 'bn_act_fwd'))
        self._bns.append(bn)

    def _bn_bwd(self, s, dout, dout2=None):
        """Stand-alone backward of a BatchNorm spec: s['dx'] from dout (+ dout2; a callable is resolved when the tape runs)."""
        c, bn, x, dx, kind = self.c, s['bn'], s['x'], s['dx'], s['kind']

        def back():
            d2 = dout2() if callable(dout2) else dout2
            _lib.check(c.L.escgnn_bn_act_bwd(
                _p(x), x.stride(0), _p(dout), dout.stride(0), _p(d2), d2.stride(0) if d2 is not None else 0, _p(s['mean']), _p(s['rstd']),
                _p(bn.weight), _p(bn.bias), ACT[s['act']], 1, _p(c.partial), _p(c.rows[kind]), c.caps[kind], x.size(1),
                _p(bn.weight.grad), _p(bn.bias.grad), _p(dx), dx.stride(0), c.st()), 'bn_act_bwd')
        self.bwd.append(back)

    def _linear(self, x, lin, kind, out=None, dx=None, dx_accumulate=False, need_dx=True, feeds_bn=False, branch=False,
                fwd_bn=None, bwd_bn=None, dy=None):
        """y = x W^T + b over capacity rows.  Returns (y, dy) buffers; backward fills W.grad, b.grad and dx.

        feeds_bn: the output goes straight into a training-mode BatchNorm.  Its bias gradient is sum_rows(dBN/dx), which
        is identically zero (BN subtracts the batch mean), so the column-sum kernels are skipped and b.grad stays at the
        zero the step starts from; torch computes the same quantity as ~1e-9 rounding noise.
        branch: nothing downstream on the critical path needs this layer's output immediately (the conv.lin edge
        projections: every layer's depends only on z): forward and dgrad run on the side stream as well; the caller waits on
        the returned events.
        fwd_bn: spec (_bn_spec) of the BatchNorm + activation that FOLLOWS this Linear: Linear, statistics, normalisation and
        activation are one launch (GEMM epilogue behind a grid barrier, csrc/gemm_tf32x3.cu EPI 1); the spec's x / dx become
        this Linear's y / dy.  Shapes the fused kernel cannot take run as GEMM + stand-alone BatchNorm.
        bwd_bn: spec of the BatchNorm + activation that PRECEDES this Linear (its `out` is our x): the dgrad's epilogue applies
        the activation / BatchNorm backward and writes the gradient wrt the spec's pre-BN x (EPI 2) -- `dx` is not needed."""
        c = self.c
        W, bvec = lin.weight, lin.bias
        n_out, k_in, rows = W.size(0), W.size(1), c.caps[kind]
        y = out if out is not None else c.buf(kind, n_out)
        dy = dy if dy is not None else c.buf(kind, n_out)
        fuse = self.fuse_bn and not branch
        fuse_f = fwd_bn is not None and fuse and self._tc_ok(x, W) and c.L.escgnn_linear_bn_fusable(rows, n_out, k_in) == 1
        fuse_b = bwd_bn is not None and fuse and need_dx and not dx_accumulate and self._tc_ok(dy, W) and \
            c.L.escgnn_linear_bn_fusable(rows, k_in, n_out) == 1
        if fwd_bn is not None:
            fwd_bn['x'], fwd_bn['dx'] = y, dy
            feeds_bn = True
        # forward: Y[rows, n_out] = X[rows, k_in] W[n_out, k_in]^T + b          (A, B K-major)
        fwd_gemm = lambda: self._gemm('gemm_fwd', x, False, W, False, y, bvec, rows, n_out, k_in, False, rows=kind)
        fwd_event = [None]
        if fuse_f:
            bn, o = fwd_bn['bn'], fwd_bn['out']
            self.fwd.append(lambda: _lib.check(c.L.escgnn_linear_bn_act_fwd(
                _p(x), x.stride(0), _p(W), W.stride(0), _p(bvec), rows, n_out, k_in, _p(c.rows[kind]), _p(bn.weight), _p(bn.bias),
                _p(bn.running_mean), _p(bn.running_var), _p(fwd_bn['mean']), _p(fwd_bn['rstd']), ACT[fwd_bn['act']], bn.eps, bn.momentum,
                _p(y), y.stride(0), _p(o), o.stride(0), _p(self.bn_ws), self.bn_ws.numel(), c.st()), 'linear_bn_act_fwd'))
            self._bns.append(bn)
        elif branch:
            self.fwd.append(lambda: fwd_event.__setitem__(0, self._fork(fwd_gemm)))
        else:
            self.fwd.append(fwd_gemm)
        if fwd_bn is not None and not fuse_f:
            self._bn_fwd(fwd_bn)
        if bwd_bn is not None and not fuse_b:
            if dx is None:
                dx = c.buf(kind, k_in)
            self._bn_bwd(bwd_bn, dx)                 # registered first: runs right AFTER this Linear's dgrad

        def grads():
            # wgrad: dW[n_out, k_in] = dY^T X   (A = dY stored [rows, n_out] = MN-major, B = X stored [rows, k_in] = MN-major)
            self._gemm('gemm_wgrad', dy, True, x, True, W.grad, None, n_out, k_in, rows, self.wgrad_mode, rows=kind)
            if not feeds_bn:
                _lib.check(c.L.escgnn_colsum(_p(dy), dy.stride(0), _p(c.rows[kind]), c.caps[kind], dy.size(1),
                                             _p(self.side_partial), _p(bvec.grad), c.st()), 'colsum')

        def dgrad():
            # dgrad: dX[rows, k_in] = dY W      (A = dY K-major, B = W stored [n_out, k_in] = MN-major for this product)
            if fuse_b:
                b = bwd_bn
                bn, xb, dxb = b['bn'], b['x'], b['dx']
                _lib.check(c.L.escgnn_linear_bn_act_bwd(
                    _p(dy), dy.stride(0), _p(W), W.stride(0), rows, k_in, n_out, _p(c.rows[kind]), _p(xb), xb.stride(0), _p(b['mean']),
                    _p(b['rstd']), _p(bn.weight), _p(bn.bias), ACT[b['act']], bn.num_features, _p(bn.weight.grad), _p(bn.bias.grad),
                    _p(dxb), dxb.stride(0), _p(self.bn_ws), self.bn_ws.numel(), c.st()), 'linear_bn_act_bwd')
                return
            self._gemm('gemm_dgrad', dy, False, W, True, dx, None, rows, k_in, n_out, dx_accumulate, rows=kind)

        def back():
            if branch:
                self._fork(lambda: (grads(), dgrad() if need_dx else None))
            else:
                self._fork(grads)                   # weight / bias gradients: only Adam consumes them
                if need_dx:
                    dgrad()
        self.bwd.append(back)
        if branch:
            return y, dy, fwd_event
        return y, dy

    def _bn_act(self, x, dx, bn, act, kind, out, dout, dout2=None, use_bn=True):
        """out = act(BN(x)) (training mode).  Backward: dx from dout (+ dout2)."""
        c = self.c
        C = x.size(1)
        if not use_bn:
            self.fwd.append(lambda: _lib.check(c.L.escgnn_act_fwd(_p(x), x.stride(0), ACT[act], _p(c.rows[kind]),
                                                                  c.caps[kind], C, _p(out), out.stride(0), c.st()), 'act_fwd'))
            self.bwd.append(lambda: _lib.check(c.L.escgnn_act_bwd(_p(x), x.stride(0), _p(dout), dout.stride(0), ACT[act],
                                                                  _p(c.rows[kind]), c.caps[kind], C, _p(dx), dx.stride(0),
                                                                  c.st()), 'act_bwd'))
            return
        mean = torch.zeros(C, dtype=torch.float32, device=c.dev)
        rstd = torch.ones(C, dtype=torch.float32, device=c.dev)
        self.fwd.append(lambda: _lib.check(c.L.escgnn_bn_act_fwd(
            _p(x), x.stride(0), _p(bn.weight), _p(bn.bias), _p(bn.running_mean), _p(bn.running_var), _p(mean), _p(rstd),
            _p(c.partial), ACT[act], bn.eps, bn.momentum, 1, _p(c.rows[kind]), c.caps[kind], C, _p(out), out.stride(0),
            c.st()), 'bn_act_fwd'))
        self._bns.append(bn)
        self.bwd.append(lambda: _lib.check(c.L.escgnn_bn_act_bwd(
            _p(x), x.stride(0), _p(dout), dout.stride(0), _p(dout2), dout2.stride(0) if dout2 is not None else 0, _p(mean),
            _p(rstd), _p(bn.weight), _p(bn.bias), ACT[act], 1, _p(c.partial), _p(c.rows[kind]), c.caps[kind], C,
            _p(bn.weight.grad), _p(bn.bias.grad), _p(dx), dx.stride(0), c.st()), 'bn_act_bwd'))

    def _embedding(self, table, idx, kind, out, side=False):
        """side: the first consumer is ordered behind a LATER fork of the side branch (the grouped projection GEMM, whose
        completion the first GINE layer waits for), so the lookup leaves the critical path."""
        c = self.c
        C = table.weight.size(1)
        cols = 1 if idx.dim() == 1 else idx.size(1)
        run = lambda: _lib.check(c.L.escgnn_embedding_fwd(_p(table.weight), _p(idx), cols, None, _p(c.rows[kind]),
                                                          c.caps[kind], C, _p(out), out.stride(0), c.st()), 'embedding_fwd')
        if side:
            ready = [None]
            self._emb_ready.append(ready)             # the first consumer on the main branch waits for these
            self.fwd.append(lambda: ready.__setitem__(0, self._fork(run)))
        else:
            self.fwd.append(run)
        return C

    def _embedding_bwd(self, table, idx, kind, dout):
        """Table gradients are read by the optimiser only: side branch."""
        c = self.c
        C = table.weight.size(1)
        cols = 1 if idx.dim() == 1 else idx.size(1)
        self.bwd.append(lambda: self._fork(lambda: _lib.check(c.L.escgnn_embedding_bwd_small(
            _p(dout), dout.stride(0), _p(idx), cols, None, _p(c.rows[kind]), c.caps[kind], C, table.weight.size(0),
            _p(table.weight.grad), c.st()), 'embedding_bwd')))

    def _gine(self, x, dx, ee, dee, eps, out, dout):
        """out = (1+eps) x + sum relu(x_src + ee).  x / ee / dee may be column slices (leading dimensions are passed).
        dee rows past the edge count are kept zero by the caller (one memset of the whole projection buffer per step)."""
        c = self.c
        C = x.size(1)
        dots = torch.zeros(c.caps['N'], dtype=torch.float32, device=c.dev)
        self.fwd.append(lambda: _lib.check(c.L.escgnn_gine_aggregate_fwd_ld(
            _p(x), x.stride(0), _p(ee), ee.stride(0), _p(self.ei[0]), _p(self.dst_ptr), _p(self.dst_perm), _p(eps), c.caps['N'], C,
            _p(out), out.stride(0), _p(c.rows['N']), c.st()), 'gine_aggregate_fwd_ld'))
        # (the backward tape runs in REVERSE order of registration: the aggregation's backward first, then the fork)
        # eps.grad = sum of the per-node dot products: only the optimiser reads it, so it leaves the critical path
        self.bwd.append(lambda: self._fork(lambda: _lib.check(c.L.escgnn_reduce_sum(_p(dots), c.caps['N'], _p(eps.grad), 0, c.st()),
                                                              'reduce_sum')))
        self.bwd.append(lambda: _lib.check(c.L.escgnn_gine_aggregate_bwd_ld(
            _p(dout), dout.stride(0), _p(x), x.stride(0), _p(ee), ee.stride(0), _p(self.ei[1]), _p(self.src_ptr),
            _p(self.src_perm), _p(eps), c.caps['N'], C, _p(dx), dx.stride(0), _p(dee), _p(dots), None, _p(c.rows['N']),
            c.st()), 'gine_aggregate_bwd_ld_noeps'))

    # ------------------------------------------------------------------ model tape
    def _build_model_tape(self):
        m, c = self.model, self.c
        H = m.z_initial.weight.size(1)
        act = 'elu' if self.variant == 'zinc' else 'relu'
        convs = [m.conv1] + list(m.convs)
        Lh = len(convs)
        # node input
        if self.variant == 'zinc':
            x0 = c.buf('N', 32)
            self._embedding(m.node_type_embedding, self.in_x, 'N', x0, side=True)
            dx0 = c.buf('N', 32)
            self._embedding_bwd(m.node_type_embedding, self.in_x, 'N', dx0)
            edge_dim = H + 32
        else:
            x0, dx0, edge_dim = self.in_x, c.buf('N', 10), H
        # M1 bag-embed + M2 z_embedding
        zcat, dzcat = c.buf('E', edge_dim), c.buf('E', edge_dim)
        if self.variant == 'zinc':                  # lookups first: their side-branch forks queue ahead of the record transposition
            self._embedding(m.edge_type_embedding, self.in_ea, 'E', zcat[:, H:], side=True)
        z0, dz0 = c.buf('E', H), c.buf('E', H)
        W0 = m.z_initial.weight
        self.fwd.append(lambda: _lib.check(c.L.escgnn_bag_embed_fwd(_p(W0), H, None, None, None, _p(self.rec), _p(self.rec_off),
                                                                    _p(self.rec_nnz), c.caps['E'], _p(z0), _p(c.rows['E']),
                                                                    c.st()), 'bag_embed_fwd'))
        bag_work = torch.zeros(3 * 1800 + 8, dtype=torch.int32, device=c.dev)
        bag_edge = torch.zeros(c.caps['nnz'], dtype=torch.int32, device=c.dev)
        bag_cnt = torch.zeros(c.caps['nnz'], dtype=torch.float32, device=c.dev)
        # the index-major transposition of the records depends only on the encoding: a side branch builds it while the
        # forward pass runs, the backward pass (its very last kernel) only reduces
        bag_ready = [None]
        self.fwd.append(lambda: bag_ready.__setitem__(0, self._fork(lambda: _lib.check(c.L.escgnn_bag_index_build(
            _p(self.rec), _p(self.rec_off), _p(self.rec_nnz), c.caps['E'], _p(bag_work), _p(bag_edge), _p(bag_cnt),
            _p(c.rows['E']), c.st()), 'bag_index_build'))))
        self.bwd.append(lambda: (torch.cuda.current_stream(c.dev).wait_event(bag_ready[0]), _lib.check(
            c.L.escgnn_bag_embed_bwd_indexed(_p(dz0), H, c.caps['nnz'], _p(W0.grad), _p(bag_work), _p(bag_edge), _p(bag_cnt),
                                             c.st()), 'bag_embed_bwd_indexed')))
        # z_embedding = BN, act, Linear, BN, act (zinc_models.py:513-522).  Forward: the first BatchNorm stand-alone, Linear + second
        # BatchNorm + activation one launch.  Backward: the second BatchNorm's is the epilogue of the grouped projection dgrad below
        # (which writes d z2 into the first H columns of `dzcat`), the first one's the epilogue of this Linear's dgrad.
        z1 = c.buf('E', H)
        bn_z1 = self._bn_spec(m.z_embedding[1], act, 'E', x=z0, dx=dz0, out=z1)
        self._bn_fwd(bn_z1)
        bn_z2 = self._bn_spec(m.z_embedding[5], act, 'E', out=zcat[:, :H])
        self._linear(z1, m.z_embedding[3], 'E', fwd_bn=bn_z2, bwd_bn=bn_z1, dy=dzcat[:, :H])
        if self.variant == 'zinc':
            self._embedding_bwd(m.edge_type_embedding, self.in_ea, 'E', dzcat[:, H:])
        # JK buffer: [x_embedding(x) | x1 .. xL] for count, [x1 .. xL] for zinc
        jk_slots = Lh + (1 if self.variant == 'count' else 0)
        xs, dxs = c.buf('N', jk_slots * H), c.buf('N', jk_slots * H)
        slot0 = 0
        if self.variant == 'count':       # xs[0] = x_embedding(data.x)   (run_graphcount.py:166)
            seq = m.x_embedding
            dxin = c.buf('N', 10)
            b_ = c.buf('N', H)
            bn_a = self._bn_spec(seq[2], act, 'N', out=b_)
            self._linear(x0, seq[0], 'N', dx=dxin, need_dx=False, fwd_bn=bn_a)
            bn_d = self._bn_spec(seq[6], act, 'N', out=xs[:, 0:H])
            self._linear(b_, seq[4], 'N', fwd_bn=bn_d, bwd_bn=bn_a)
            self._bn_bwd(bn_d, dxs[:, 0:H])            # registered last of the three: runs first in the backward pass
            slot0 = 1
        # M3 GINE layers
        x_prev, dx_prev = x0, dx0
        # ---- edge projections of every layer in one GEMM: ee_all[E, sum C_in] = zcat @ W_cat^T + b_cat
        col, off = {}, 0
        for cv in self.lin_convs:
            col[id(cv)] = off
            off += cv.lin.weight.size(0)
        n_tot, ld_all = off, (off + 3) // 4 * 4
        flat, gflat = self.opt.flat, self.opt.grad
        w0 = (self.lin_convs[0].lin.weight.data_ptr() - flat.data_ptr()) // 4
        b0 = (self.lin_convs[0].lin.bias.data_ptr() - flat.data_ptr()) // 4
        W_cat, dW_cat = flat[w0:w0 + n_tot * edge_dim].view(n_tot, edge_dim), gflat[w0:w0 + n_tot * edge_dim].view(n_tot, edge_dim)
        b_cat, db_cat = flat[b0:b0 + n_tot], gflat[b0:b0 + n_tot]
        assert self.lin_convs[-1].lin.weight.data_ptr() == W_cat[col[id(self.lin_convs[-1])]].data_ptr(), 'lin weights not adjacent'
        assert self.lin_convs[-1].lin.bias.data_ptr() == b_cat[col[id(self.lin_convs[-1])]:].data_ptr(), 'lin biases not adjacent'
        ee_all, dee_all = c.buf('E', ld_all), c.buf('E', ld_all)
        E_rows = c.caps['E']
        ee_ready = [None]
        # rows past the edge count stay zero for the GEMMs of the backward pass (a larger earlier batch may have written them);
        # side branch, ahead of the projection GEMM whose completion the first GINE layer waits for
        self.fwd.append(lambda: self._fork(lambda: _lib.check(c.L.escgnn_zero_tail_rows(
            _p(dee_all), dee_all.stride(0), n_tot, _p(c.rows['E']), E_rows, c.st()), 'zero_tail_rows')))
        # forward projections: the first layer's narrow block (the LAST rows of W_cat) on the main branch, right where it is
        # needed; the wide blocks of layers 2..L on the side branch -- the second GINE layer is the first to wait for them
        c1 = col[id(self.lin_convs[-1])]
        n_rest = c1                                   # columns of layers 2..L
        if n_rest > 0:
            self.fwd.append(lambda: ee_ready.__setitem__(0, self._fork(
                lambda: self._gemm('gemm_fwd', zcat, False, W_cat[:n_rest], False, ee_all[:, :n_rest], b_cat[:n_rest], E_rows, n_rest,
                                   edge_dim, False, rows='E'))))

        def first_projection():
            main = torch.cuda.current_stream(c.dev)
            for ready in self._emb_ready:             # x0 and the edge-type columns of zcat come from the side branch
                main.wait_event(ready[0])
            self._gemm('gemm_fwd', zcat, False, W_cat[c1:], False, ee_all[:, c1:n_tot], b_cat[c1:], E_rows, n_tot - c1, edge_dim,
                       False, rows='E')
        self.fwd.append(first_projection)

        # d zcat = dee_all W_cat.  Its first H columns are the output of z_embedding's last BatchNorm + activation: that BatchNorm's
        # backward is the epilogue of this product (-> d z2 lands in dzcat[:, :H]); the edge-type columns (ZINC) only feed an
        # embedding-table gradient, so their (narrow) product leaves the critical path
        fuse_proj = self.fuse_bn and self._tc_ok(dee_all, W_cat) and c.L.escgnn_linear_bn_fusable(E_rows, H, n_tot) == 1
        split_proj = self.split_proj and not fuse_proj
        proj_done = [None]                        # completion of the latest per-layer term (they run in order on one branch)
        if not fuse_proj:
            dz_act = c.buf('E', H)                # gradient wrt the activation output, BatchNorm backward as its own launch
            self._bn_bwd(bn_z2, dz_act)           # (registered before proj_back: runs after it)

        def proj_back():                          # runs after every layer's backward has filled its slice of dee_all
            def side():
                self._gemm('gemm_wgrad', dee_all, True, zcat, True, dW_cat, None, n_tot, edge_dim, E_rows, self.wgrad_mode, rows='E')
                _lib.check(c.L.escgnn_colsum(_p(dee_all), dee_all.stride(0), _p(c.rows['E']), E_rows, n_tot,
                                             _p(self.side_partial), _p(db_cat), c.st()), 'colsum')
                if edge_dim > H:
                    self._gemm('gemm_dgrad', dee_all, False, W_cat[:, H:], True, dzcat[:, H:], None, E_rows, edge_dim - H, n_tot, False, rows='E')
            self._fork(side)
            if fuse_proj:
                bn, b = bn_z2['bn'], bn_z2
                _lib.check(c.L.escgnn_linear_bn_act_bwd(
                    _p(dee_all), dee_all.stride(0), _p(W_cat), W_cat.stride(0), E_rows, H, n_tot, _p(c.rows['E']), _p(b['x']), b['x'].stride(0),
                    _p(b['mean']), _p(b['rstd']), _p(bn.weight), _p(bn.bias), ACT[b['act']], H, _p(bn.weight.grad), _p(bn.bias.grad),
                    _p(dzcat), dzcat.stride(0), _p(self.bn_ws), self.bn_ws.numel(), c.st()), 'linear_bn_act_bwd')
            elif split_proj:                      # every layer's term is already on its way: wait for the last one
                torch.cuda.current_stream(c.dev).wait_event(proj_done[0])
            else:
                self._gemm('gemm_dgrad', dee_all, False, W_cat[:, :H], True, dz_act, None, E_rows, H, n_tot, False, rows='E')
            # every gradient outside the late bucket is complete once the side branch has passed this point, and nothing below
            # reads the parameters of the early bucket any more (this dgrad was the last reader of W_cat): exchange + update them
            # on the side branch, under the z path's backward
            if self.exchange == 'p2p' and self.bucketed_exchange and not self.inline_branches and self.opt.tail_begin > 0:
                main = torch.cuda.current_stream(c.dev)             # its own branch: behind this dgrad AND behind the side branch's
                ev_m, ev_s = torch.cuda.Event(), torch.cuda.Event()  # weight gradients, but not in front of the side work still to come
                ev_m.record(main)
                ev_s.record(self.side)
                with torch.cuda.stream(self.xchg):
                    self.xchg.wait_event(ev_m)
                    self.xchg.wait_event(ev_s)
                    self.opt.step_exchange_device(0, self.opt.tail_begin, bucket=0, tick=True)
                    self._xchg_done = torch.cuda.Event()
                    self._xchg_done.record(self.xchg)
                self._early_done = True
        self.bwd.append(proj_back)
        layer_dx_from_next = [None] * Lh          # gradient flowing into layer l's output from layer l+1's aggregation
        for l, conv in enumerate(convs):
            cin = conv.lin.weight.size(0)
            c0 = col[id(conv)]
            ee, dee = ee_all[:, c0:c0 + cin], dee_all[:, c0:c0 + cin]
            agg, dagg = c.buf('N', cin), c.buf('N', cin)
            if l == 0:
                xin, dxin_buf = x_prev, dx_prev
            else:                                  # the previous layer's output is a column slice of the JK buffer
                if l == 1:
                    self.fwd.append(lambda: torch.cuda.current_stream(c.dev).wait_event(ee_ready[0]))
                xin = xs[:, (slot0 + l - 1) * H:(slot0 + l) * H]
                dxin_buf = c.buf('N', H)
                layer_dx_from_next[l - 1] = dxin_buf
            if split_proj:
                # (registered before the aggregation: runs right after its backward)  d z (+)= dee_l W_l[:, :H]; the last layer's
                # backward runs first and overwrites, the others accumulate
                def proj_term(l=l, dee=dee, c0=c0, cin=cin):
                    proj_done[0] = self._fork(lambda: self._gemm('gemm_dgrad', dee, False, W_cat[c0:c0 + cin, :H], True, dz_act, None,
                                                                 E_rows, H, cin, l != Lh - 1, rows='E'), stream=self.proj_stream)
                self.bwd.append(proj_term)
            self._gine(xin, dxin_buf, ee, dee, conv.eps, agg, dagg)
            # conv.nn = Linear, BN, act, Linear, BN, act: two fused launches forward; backward: the last BatchNorm stand-alone (its
            # gradient arrives from two places), the middle one as the epilogue of the second Linear's dgrad
            seq = conv.nn
            h2 = c.buf('N', H)
            out_slice = xs[:, (slot0 + l) * H:(slot0 + l + 1) * H]
            dout_slice = dxs[:, (slot0 + l) * H:(slot0 + l + 1) * H]
            bn_mid = self._bn_spec(seq[2], act, 'N', out=h2)
            bn_out = self._bn_spec(seq[6], act, 'N', out=out_slice)
            self._linear(agg, seq[0], 'N', dx=dagg, fwd_bn=bn_mid)
            self._linear(h2, seq[4], 'N', fwd_bn=bn_out, bwd_bn=bn_mid)
            # the gradient of layer l's output also comes from layer l+1's aggregation, known only after the loop wiring: resolved
            # when the tape runs
            self._bn_bwd(bn_out, dout_slice, dout2=lambda l=l: layer_dx_from_next[l])
        # M4 readout
        if self.variant == 'zinc':
            pooled, dpooled = c.buf('B', Lh * H), c.buf('B', Lh * H)
            self.fwd.append(lambda: _lib.check(c.L.escgnn_segment_pool_fwd(_p(xs), _p(self.graph_ptr), self.G, Lh * H, 0,
                                                                           _p(pooled), c.st()), 'segment_pool_fwd'))
            self.bwd.append(lambda: _lib.check(c.L.escgnn_segment_pool_bwd(_p(dpooled), _p(self.graph_ptr), self.G, Lh * H, 0,
                                                                           _p(dxs), c.st()), 'segment_pool_bwd'))
            head_in, dhead_in, kind = pooled, dpooled, 'B'
        else:
            head_in, dhead_in, kind = xs, dxs, 'N'
        head_bn = self.G > 1 or kind == 'N'
        p1, dp1 = self._linear(head_in, m.lin1, kind, dx=dhead_in, feeds_bn=head_bn)
        if self.fused_head and self.variant == 'zinc' and head_bn and H <= 256 and self.G <= 512:
            # BatchNorm + activation + lin2 + L1 loss AND their backward in one launch (five dependent launches otherwise);
            # dp1 is ready when the backward tape starts, so the head registers no backward entries of its own
            bn, lin2 = m.bn_lin1, m.lin2
            pred = c.buf(kind, 1)
            self.pred = pred
            self._bns.append(bn)
            self.debug_buffers = dict(head_in=head_in, dhead_in=dhead_in, p1=p1, dp1=dp1, pred=pred, xs=xs, dxs=dxs, zcat=zcat,
                                      dzcat=dzcat)
            self.fwd.append(lambda: _lib.check(c.L.escgnn_head_bn_linear_l1(
                _p(p1), p1.stride(0), _p(bn.weight), _p(bn.bias), _p(bn.running_mean), _p(bn.running_var), ACT[act], bn.eps,
                bn.momentum, _p(lin2.weight), _p(lin2.bias), _p(self.in_y), _p(c.rows[kind]), c.caps[kind], H, _p(pred),
                _p(self.loss), _p(dp1), dp1.stride(0), _p(bn.weight.grad), _p(bn.bias.grad), _p(lin2.weight.grad),
                _p(lin2.bias.grad), c.st()), 'head_bn_linear_l1'))
            return
        p2, dp2 = c.buf(kind, H), c.buf(kind, H)
        self._bn_act(p1, dp1, m.bn_lin1, act, kind, p2, dp2, use_bn=head_bn)
        pred, dpred = self._linear(p2, m.lin2, kind, dx=dp2)
        self.pred = pred
        self.debug_buffers = dict(head_in=head_in, dhead_in=dhead_in, p1=p1, dp1=dp1, p2=p2, dp2=dp2, pred=pred, dpred=dpred,
                                  xs=xs, dxs=dxs, zcat=zcat, dzcat=dzcat)
        self.fwd.append(lambda: _lib.check(c.L.escgnn_loss_fwd_bwd(_p(pred), pred.stride(0), _p(self.in_y), 0, _p(c.rows[kind]),
                                                                   c.caps[kind], 1, _p(self.loss), _p(dpred), dpred.stride(0),
                                                                   c.st()), 'loss_fwd_bwd'))

    # ------------------------------------------------------------------ OGB variant (ogb_mol_gnn.py:66-117,614-792)
    def _dropout(self, x, dx, kind, p):
        """(out, dout) with out = dropout(x) in training mode; the backward applies the same counter-based mask to dout.
        p == 0: no kernels, the buffers are passed through."""
        if p <= 0.0:
            return x, dx
        c = self.c
        C = x.size(1)
        out, dout = c.buf(kind, C), c.buf(kind, C)
        self._salt += 1
        salt = self._salt
        step = self.opt.state
        self.fwd.append(lambda: _lib.check(c.L.escgnn_dropout(_p(x), x.stride(0), p, salt, _p(step), _p(c.rows[kind]), c.caps[kind], C,
                                                              _p(out), out.stride(0), c.st()), 'dropout'))
        self.bwd.append(lambda: _lib.check(c.L.escgnn_dropout(_p(dout), dout.stride(0), p, salt, _p(step), _p(c.rows[kind]),
                                                              c.caps[kind], C, _p(dx), dx.stride(0), c.st()), 'dropout'))
        return out, dout

    def _build_ogb_tape(self):
        """GNN(gnn_type='gin_eff'): AtomEncoder, z path, L x [virtual-node broadcast, GINConv_eff, BN(+ReLU), dropout, virtual-node
        update], mean / sum pooling, Linear head, BCE-with-logits over labelled targets.  Registration order = a topological order
        of the forward pass; the backward tape is its reverse, so for every buffer with several consumers the LAST registered
        consumer's backward writes the gradient and the earlier ones accumulate."""
        m, c = self.model, self.c
        gn = m.gnn_node
        H, Lh, p = m.emb_dim, gn.num_layer, float(gn.drop_ratio)
        G = self.G
        self._salt = 0
        i64 = torch.int64
        dev = c.dev
        # ---- node input: sum of the 9 atom tables (adjacent in the flat buffer -> one lookup with row offsets)
        atom = list(gn.node_encoder.atom_embedding_list)
        atom_off = torch.tensor([0] + list(torch.tensor([e.weight.size(0) for e in atom]).cumsum(0)[:-1]), dtype=i64, device=dev)
        h = [c.buf('N', H) for _ in range(Lh + 1)]
        dh = [c.buf('N', H) for _ in range(Lh + 1)]
        a0, ga0 = atom[0].weight, atom[0].weight.grad
        assert atom[-1].weight.data_ptr() == a0.data_ptr() + 4 * H * int(atom_off[-1]), 'atom tables not adjacent'
        self.fwd.append(lambda: _lib.check(c.L.escgnn_embedding_fwd(_p(a0), _p(self.in_x), 9, _p(atom_off), _p(c.rows['N']), c.caps['N'],
                                                                    H, _p(h[0]), H, c.st()), 'embedding_fwd'))
        self.bwd.append(lambda: self._fork(lambda: _lib.check(c.L.escgnn_embedding_bwd(
            _p(dh[0]), H, _p(self.in_x), 9, _p(atom_off), _p(c.rows['N']), c.caps['N'], H, _p(ga0), c.st()), 'embedding_bwd')))
        # ---- z path: bag-embed -> Dropout, BN, ReLU, Linear, Dropout, BN, ReLU
        z0, dz0 = c.buf('E', H), c.buf('E', H)
        W0 = gn.z_initial.weight
        self.fwd.append(lambda: _lib.check(c.L.escgnn_bag_embed_fwd(_p(W0), H, None, None, None, _p(self.rec), _p(self.rec_off),
                                                                    _p(self.rec_nnz), c.caps['E'], _p(z0), _p(c.rows['E']),
                                                                    c.st()), 'bag_embed_fwd'))
        bag_work = torch.zeros(3 * 1800 + 8, dtype=torch.int32, device=dev)
        bag_edge = torch.zeros(c.caps['nnz'], dtype=torch.int32, device=dev)
        bag_cnt = torch.zeros(c.caps['nnz'], dtype=torch.float32, device=dev)
        bag_ready = [None]
        self.fwd.append(lambda: bag_ready.__setitem__(0, self._fork(lambda: _lib.check(c.L.escgnn_bag_index_build(
            _p(self.rec), _p(self.rec_off), _p(self.rec_nnz), c.caps['E'], _p(bag_work), _p(bag_edge), _p(bag_cnt),
            _p(c.rows['E']), c.st()), 'bag_index_build'))))
        self.bwd.append(lambda: (torch.cuda.current_stream(dev).wait_event(bag_ready[0]), _lib.check(
            c.L.escgnn_bag_embed_bwd_indexed(_p(dz0), H, c.caps['nnz'], _p(W0.grad), _p(bag_work), _p(bag_edge), _p(bag_cnt),
                                             c.st()), 'bag_embed_bwd_indexed')))
        ze = gn.z_embedding
        z0d, dz0d = self._dropout(z0, dz0, 'E', p)
        z1, dz1 = c.buf('E', H), c.buf('E', H)
        self._bn_act(z0d, dz0d, ze[1], 'relu', 'E', z1, dz1)
        z2, dz2 = self._linear(z1, ze[3], 'E', dx=dz1, feeds_bn=(p <= 0.0))
        z2d, dz2d = self._dropout(z2, dz2, 'E', p)
        zemb, dzemb = c.buf('E', H), c.buf('E', H)
        self._bn_act(z2d, dz2d, ze[5], 'relu', 'E', zemb, dzemb)
        # ---- edge features of every layer: e_l = BondEncoder_l(edge_attr) + Linear_l(z): the lookups write the slices of one
        # [E, L*H] buffer, then ONE grouped GEMM accumulates all projections on top (C += zemb W_cat^T + b_cat)
        convs = list(gn.convs)
        flat, gflat = self.opt.flat, self.opt.grad
        w0 = (convs[0].edge_encoder_pos.weight.data_ptr() - flat.data_ptr()) // 4
        b0 = (convs[0].edge_encoder_pos.bias.data_ptr() - flat.data_ptr()) // 4
        n_tot = Lh * H
        W_cat, dW_cat = flat[w0:w0 + n_tot * H].view(n_tot, H), gflat[w0:w0 + n_tot * H].view(n_tot, H)
        b_cat, db_cat = flat[b0:b0 + n_tot], gflat[b0:b0 + n_tot]
        assert convs[-1].edge_encoder_pos.weight.data_ptr() == W_cat[(Lh - 1) * H].data_ptr(), 'projection weights not adjacent'
        assert convs[-1].edge_encoder_pos.bias.data_ptr() == b_cat[(Lh - 1) * H:].data_ptr(), 'projection biases not adjacent'
        ld_all = (n_tot + 3) // 4 * 4
        e_all, de_all = c.buf('E', ld_all), c.buf('E', ld_all)
        E_rows = c.caps['E']
        bond_off = []
        for l, cv in enumerate(convs):
            tabs = list(cv.edge_encoder.bond_embedding_list)
            off = torch.tensor([0] + list(torch.tensor([e.weight.size(0) for e in tabs]).cumsum(0)[:-1]), dtype=i64, device=dev)
            assert tabs[-1].weight.data_ptr() == tabs[0].weight.data_ptr() + 4 * H * int(off[-1]), 'bond tables not adjacent'
            bond_off.append(off)
            t0, e_l, de_l = tabs[0].weight, e_all[:, l * H:(l + 1) * H], de_all[:, l * H:(l + 1) * H]
            self.fwd.append(lambda t0=t0, off=off, e_l=e_l: _lib.check(c.L.escgnn_embedding_fwd(
                _p(t0), _p(self.in_ea), 3, _p(off), _p(c.rows['E']), E_rows, H, _p(e_l), e_l.stride(0), c.st()), 'embedding_fwd'))
            n_bond = sum(e.weight.size(0) for e in tabs)
            self.bwd.append(lambda t0=t0, off=off, de_l=de_l, n_bond=n_bond: self._fork(lambda: _lib.check(
                c.L.escgnn_embedding_bwd_small(_p(de_l), de_l.stride(0), _p(self.in_ea), 3, _p(off), _p(c.rows['E']), E_rows, H, n_bond,
                                               _p(t0.grad), c.st()), 'embedding_bwd')))
        self.fwd.append(lambda: _lib.check(c.L.escgnn_zero_tail_rows(_p(de_all), de_all.stride(0), n_tot, _p(c.rows['E']), E_rows,
                                                                     c.st()), 'zero_tail_rows'))
        self.fwd.append(lambda: self._gemm('gemm_fwd', zemb, False, W_cat, False, e_all, b_cat, E_rows, n_tot, H, True, rows='E'))

        def proj_back():                          # after every layer's backward: d zemb, d W_cat, d b_cat from de_all
            self._fork(lambda: (self._gemm('gemm_wgrad', de_all, True, zemb, True, dW_cat, None, n_tot, H, E_rows, self.wgrad_mode,
                                           rows='E'),
                                _lib.check(c.L.escgnn_colsum(_p(de_all), de_all.stride(0), _p(c.rows['E']), E_rows, n_tot,
                                                             _p(self.side_partial), _p(db_cat), c.st()), 'colsum')))
            self._gemm('gemm_dgrad', de_all, False, W_cat, True, dzemb, None, E_rows, H, n_tot, False, rows='E')
        self.bwd.append(proj_back)
        # ---- virtual node
        vnode = gn.virtual_node
        if vnode:
            vn = [c.buf('B', H) for _ in range(Lh)]
            dvn = [c.buf('B', H) for _ in range(Lh)]
            zeros_idx = torch.zeros(G, dtype=i64, device=dev)
            vw = gn.virtualnode_embedding.weight
            self.fwd.append(lambda: _lib.check(c.L.escgnn_embedding_fwd(_p(vw), _p(zeros_idx), 1, None, _p(c.rows['B']), G, H,
                                                                        _p(vn[0]), H, c.st()), 'embedding_fwd'))
            self.bwd.append(lambda: self._fork(lambda: _lib.check(c.L.escgnn_colsum(
                _p(dvn[0]), H, _p(c.rows['B']), G, H, _p(self.side_partial), _p(vw.grad), c.st()), 'colsum')))
        gptr = self.graph_ptr
        for l, cv in enumerate(convs):
            last = l == Lh - 1
            e_l, de_l = e_all[:, l * H:(l + 1) * H], de_all[:, l * H:(l + 1) * H]
            if vnode:
                # F1: h_l += vn_l[batch] (in place); backward: d vn_l (+)= segment sums of d h_l.  Registered first, so it runs
                # LAST of this layer's backward entries, when d h_l is complete.
                self.fwd.append(lambda l=l: _lib.check(c.L.escgnn_add_segment_rows(_p(h[l]), H, _p(vn[l]), H, _p(gptr), G, H, _p(h[l]), H,
                                                                                   c.st()), 'add_segment_rows'))
                if last:
                    self.bwd.append(lambda l=l: _lib.check(c.L.escgnn_segment_pool_fwd(_p(dh[l]), _p(gptr), G, H, 0, _p(dvn[l]), c.st()),
                                                           'segment_pool_fwd'))
                else:
                    seg_tmp = c.buf('B', H)

                    def vn_grad(l=l, seg_tmp=seg_tmp):       # d vn_l = (dgrad of the update MLP, already there) + segment sums
                        _lib.check(c.L.escgnn_segment_pool_fwd(_p(dh[l]), _p(gptr), G, H, 0, _p(seg_tmp), c.st()), 'segment_pool_fwd')
                        dvn[l].add_(seg_tmp)
                        _lib.mark('misc')
                    self.bwd.append(vn_grad)
                    if gn.residual:                  # vn_{l+1} = vn_l + update: d vn_l += d vn_{l+1} (after the dgrad below has written)
                        self.bwd.append(lambda l=l: (dvn[l].add_(dvn[l + 1]), _lib.mark('misc')))
                    # F4: vn_{l+1} = dropout(MLP(add_pool(h_l) + vn_l)); the sum goes through the first Linear as two products
                    seq = gn.mlp_virtualnode_list[l]
                    pooled, dpooled = c.buf('B', H), c.buf('B', H)
                    self.fwd.append(lambda l=l, pooled=pooled: _lib.check(c.L.escgnn_segment_pool_fwd(
                        _p(h[l]), _p(gptr), G, H, 0, _p(pooled), c.st()), 'segment_pool_fwd'))
                    # backward of the pooling: d h_l += d pooled[batch]  (runs after the aggregation's backward has written d h_l)
                    self.bwd.append(lambda l=l, dpooled=dpooled: _lib.check(c.L.escgnn_add_segment_rows(
                        _p(dh[l]), H, _p(dpooled), H, _p(gptr), G, H, _p(dh[l]), H, c.st()), 'add_segment_rows'))
                    lin_a = seq[0]
                    t1, dt1 = self._linear(pooled, lin_a, 'B', dx=dpooled, feeds_bn=True)
                    # second product of the same Linear: t1 += vn_l W^T; backward: d vn_l = d t1 W (writes), d W += d t1^T vn_l
                    self.fwd.append(lambda l=l, t1=t1, lin_a=lin_a: self._gemm('gemm_fwd', vn[l], False, lin_a.weight, False, t1, None,
                                                                              G, 2 * H, H, True, rows='B'))

                    def vn_lin_back(l=l, dt1=dt1, lin_a=lin_a):
                        self._fork(lambda: self._gemm('gemm_wgrad', dt1, True, vn[l], True, lin_a.weight.grad, None, 2 * H, H, G,
                                                      2 if self.wgrad_mode == 2 else 1, rows='B'))
                        self._gemm('gemm_dgrad', dt1, False, lin_a.weight, True, dvn[l], None, G, H, 2 * H, False, rows='B')
                    self.bwd.append(vn_lin_back)
                    t2, dt2 = c.buf('B', 2 * H), c.buf('B', 2 * H)
                    self._bn_act(t1, dt1, seq[1], 'relu', 'B', t2, dt2)
                    t3, dt3 = self._linear(t2, seq[3], 'B', dx=dt2, feeds_bn=True)
                    t4, dt4 = c.buf('B', H), c.buf('B', H)
                    self._bn_act(t3, dt3, seq[4], 'relu', 'B', t4, dt4)
                    if p > 0.0:
                        self._salt += 1
                        salt, step = self._salt, self.opt.state
                        self.fwd.append(lambda l=l, t4=t4, salt=salt: _lib.check(c.L.escgnn_dropout(
                            _p(t4), H, p, salt, _p(step), _p(c.rows['B']), G, H, _p(vn[l + 1]), H, c.st()), 'dropout'))
                        self.bwd.append(lambda l=l, dt4=dt4, salt=salt: _lib.check(c.L.escgnn_dropout(
                            _p(dvn[l + 1]), H, p, salt, _p(step), _p(c.rows['B']), G, H, _p(dt4), H, c.st()), 'dropout'))
                    else:
                        self.fwd.append(lambda l=l, t4=t4: (vn[l + 1].copy_(t4), _lib.mark('copy')))
                        self.bwd.append(lambda l=l, dt4=dt4: (dt4.copy_(dvn[l + 1]), _lib.mark('copy')))
                    if gn.residual:
                        self.fwd.append(lambda l=l: (vn[l + 1].add_(vn[l]), _lib.mark('misc')))
            if gn.residual:      # h_{l+1} = layer(h_l) + h_l: d h_l += d h_{l+1}; registered here so that it runs AFTER the aggregation's
                self.bwd.append(lambda l=l: (dh[l].add_(dh[l + 1]), _lib.mark('misc')))      # backward has written d h_l
            # F2: GINConv_eff aggregation, F3: its MLP, the layer's BatchNorm (+ReLU except after the last layer), dropout
            agg, dagg = c.buf('N', H), c.buf('N', H)
            self._gine(h[l], dh[l], e_l, de_l, cv.eps, agg, dagg)
            m1, dm1 = self._linear(agg, cv.mlp[0], 'N', dx=dagg, feeds_bn=True)
            m2, dm2 = c.buf('N', 2 * H), c.buf('N', 2 * H)
            self._bn_act(m1, dm1, cv.mlp[1], 'relu', 'N', m2, dm2)
            m3, dm3 = self._linear(m2, cv.mlp[3], 'N', dx=dm2, feeds_bn=True)
            if p > 0.0:
                m4, dm4 = c.buf('N', H), c.buf('N', H)
                self._bn_act(m3, dm3, gn.batch_norms[l], 'none' if last else 'relu', 'N', m4, dm4)
                self._salt += 1
                salt, step = self._salt, self.opt.state
                self.fwd.append(lambda l=l, m4=m4, salt=salt: _lib.check(c.L.escgnn_dropout(
                    _p(m4), H, p, salt, _p(step), _p(c.rows['N']), c.caps['N'], H, _p(h[l + 1]), H, c.st()), 'dropout'))
                self.bwd.append(lambda l=l, dm4=dm4, salt=salt: _lib.check(c.L.escgnn_dropout(
                    _p(dh[l + 1]), H, p, salt, _p(step), _p(c.rows['N']), c.caps['N'], H, _p(dm4), H, c.st()), 'dropout'))
            else:
                self._bn_act(m3, dm3, gn.batch_norms[l], 'none' if last else 'relu', 'N', h[l + 1], dh[l + 1])
            if gn.residual:
                self.fwd.append(lambda l=l: (h[l + 1].add_(h[l]), _lib.mark('misc')))
        # ---- readout: pooling, Linear head, BCE-with-logits over labelled targets (run_ogb_mol.py:58-74)
        mean = 1 if m.graph_pooling == 'mean' else 0
        hg, dhg = c.buf('B', H), c.buf('B', H)
        self.fwd.append(lambda: _lib.check(c.L.escgnn_segment_pool_fwd(_p(h[Lh]), _p(gptr), G, H, mean, _p(hg), c.st()),
                                           'segment_pool_fwd'))
        self.bwd.append(lambda: _lib.check(c.L.escgnn_segment_pool_bwd(_p(dhg), _p(gptr), G, H, mean, _p(dh[Lh]), c.st()),
                                           'segment_pool_bwd'))
        pred, dpred = self._linear(hg, m.graph_pred_linear, 'B', dx=dhg)
        self.pred = pred
        self.debug_buffers = dict(h=h, dh=dh, pred=pred, dpred=dpred, zemb=zemb, dzemb=dzemb, e_all=e_all, de_all=de_all)
        self.fwd.append(lambda: _lib.check(c.L.escgnn_loss_fwd_bwd(_p(pred), pred.stride(0), _p(self.in_y), 1, _p(c.rows['B']), G, 1,
                                                                   _p(self.loss), _p(dpred), dpred.stride(0), c.st()),
                                           'loss_fwd_bwd'))

    # ------------------------------------------------------------------ one step
    def _encode_and_index(self, t=None):
        """E1 / E5 / E2-E4 on the raw-input buffers, device collation and index builds, written into batch set `t`."""
        t = self.live if t is None else t
        c, L, fl, G = self.c, self.c.L, self.flags, self.G
        st = c.st()
        self.counters[:_lib.CTR_PER_CALL].zero_()    # the sticky error / max-nnz slots survive (check_errors reads them)
        _lib.mark('memset')
        if fl['self_loop']:
            _lib.check(L.escgnn_rewrite_self_loops(_p(self.in_src), _p(self.in_dst), _p(self.in_eptr), _p(self.in_nptr), G,
                                                   _p(self.eo_ptr), _p(self.eo[0]), _p(self.eo[1]), _p(self.rw_tmp), st),
                       'rewrite_self_loops')
            es, ed, ep = self.eo[0], self.eo[1], self.eo_ptr
            if self.ea_via_raw:
                _lib.check(L.escgnn_rewrite_edge_attr(_p(self.in_src), _p(self.in_dst), _p(self.in_eptr), _p(self.in_nptr), G,
                                                      _p(self.eo_ptr), _p(self.in_ea_raw), self.ea_cols, 1, _p(t.in_ea), st),
                           'rewrite_edge_attr')
        else:
            es, ed, ep = self.in_src, self.in_dst, self.in_eptr
            if self.ea_via_raw:
                t.in_ea[:self.in_ea_raw.size(0)].copy_(self.in_ea_raw)
        if fl['use_rd']:
            _lib.check(L.escgnn_encode_rd(_p(es), _p(ed), _p(ep), _p(self.in_nptr), G, fl['h'], _p(self.rdh), _p(self.counters),
                                          self.max_n, self.max_e, _p(self.scratch), self.scratch.numel(), st), 'encode_rd')
        _lib.check(L.escgnn_encode(_p(es), _p(ed), _p(ep), _p(self.in_nptr), G, fl['h'], _p(self.rdh), _p(t.rec),
                                   t.rec.numel(), _p(t.rec_off), _p(t.rec_nnz), _p(self.edge_graph),
                                   _p(self.counters), self.max_n, self.max_e, _p(self.scratch), self.scratch.numel(), st),
                   'encode')
        _lib.check(L.escgnn_make_dims(_p(ep), _p(self.in_nptr), G, _p(self.counters), _p(t.dims), st), 'make_dims')
        N, E = c.caps['N'], c.caps['E']
        _lib.check(L.escgnn_collate_edges(_p(es), _p(ed), _p(self.edge_graph), _p(self.in_nptr), E, _p(t.ei[0]),
                                          _p(t.ei[1]), _p(t.rows['E']), st), 'collate_edges')
        _lib.check(L.escgnn_ptr_to_ids(_p(self.in_nptr), G, N, _p(t.batch), _p(t.rows['N']), st), 'ptr_to_ids')
        _lib.check(L.escgnn_csr_build(_p(t.ei[1]), E, N, _p(t.dst_ptr), _p(t.dst_perm), _p(t.tmp_a),
                                      _p(self.idx_err), _p(t.rows['E']), st), 'csr_build')
        _lib.check(L.escgnn_csr_build(_p(t.ei[0]), E, N, _p(t.src_ptr), _p(t.src_perm), _p(t.tmp_b),
                                      _p(self.idx_err), _p(t.rows['E']), st), 'csr_build')
        _lib.check(L.escgnn_sorted_ids_to_ptr(_p(t.batch), N, G, _p(t.graph_ptr), _p(t.rows['N']), st),
                   'sorted_ids_to_ptr')

    def _forward_features(self, t):
        """Pipelined mode: features / targets of the batch just loaded travel with the batch set they belong to."""
        t.in_x.copy_(self.raw_feat.in_x)
        t.in_y.copy_(self.raw_feat.in_y)
        if t.in_ea is not None and not self.ea_via_raw:
            t.in_ea.copy_(self.raw_feat.in_ea)

    @torch.no_grad()
    def _run_train(self):
        """Forward, loss and backward on the LIVE batch set."""
        self.opt.grad.zero_()
        _lib.mark('memset')
        for f in self.fwd:
            f()
        for b in reversed(self.bwd):
            b()
        self._join()                                 # every weight gradient is in place before the optimiser / exchange
        if self._xchg_done is not None:              # ... and so is the early bucket of the peer-memory exchange
            torch.cuda.current_stream(self.c.dev).wait_event(self._xchg_done)
            self._xchg_done = None

    @torch.no_grad()
    def _run_main(self):
        """Everything up to the gradients as a fixed launch sequence (run eagerly, or captured once and replayed)."""
        _lib.mark('start')
        if self.pipeline and not self.inline_branches:
            # batch k-1 goes live with one copy; the encoder then fills the staging set with batch k on its own branch
            # while the main branch trains on batch k-1
            main = torch.cuda.current_stream(self.c.dev)
            self.live.buf.copy_(self.stage.buf)
            moved = torch.cuda.Event()
            moved.record(main)
            with torch.cuda.stream(self.enc_stream):
                self.enc_stream.wait_event(moved)
                self._forward_features(self.stage)
                # the next batch's encoder has a whole step of slack: confined to a few SMs it stops competing with the
                # latency-critical chain of the current batch for registers and shared memory
                prev = self.c.L.escgnn_set_encoder_grid_cap(self.encoder_ctas) if self.encoder_ctas else None
                self._encode_and_index(self.stage)
                if prev is not None:
                    self.c.L.escgnn_set_encoder_grid_cap(prev)
                enc_done = torch.cuda.Event()
                enc_done.record(self.enc_stream)
            self._run_train()
            main.wait_event(enc_done)
        else:
            if self.pipeline:                        # single-stream profile of the same work
                self._forward_features(self.live)
                _lib.mark('copy')
            self._encode_and_index(self.live)
            self._run_train()

    @torch.no_grad()
    def _run_opt(self):
        if self.exchange == 'p2p':               # exchange + update in one launch (every rank runs the same number of steps)
            if self._early_done:                 # the early bucket went out on the side branch (joined by _run_train): the tail is left
                self.opt.step_exchange_device(self.opt.tail_begin, None, bucket=1, tick=False)
                self._early_done = False
            else:
                self.opt.step_exchange_device()
        else:
            self.opt.step_device()

    def _run(self):
        self._run_main()
        if self.exchange == 'nccl':
            self.opt.all_reduce_grads()
        self._run_opt()

    def load(self, raw):
        """Copy one RawBatch (pinned host or device) into the static input buffers (async on the current stream)."""
        e, g = raw.src.numel(), raw.num_graphs
        if g > self.G or g < 1:
            raise ValueError('engine built for at most %d graphs per step, got %d' % (self.G, g))
        if g == 1 and self.G > 1 and self.variant != 'ogb':
            # the reference skips bn_lin1 when the batch has a single row (zinc_models.py:605-606); the engine's readout tail is
            # built once from the capacity
            raise NotImplementedError('a one-graph batch needs an engine built with max_graphs=1')
        if e > self.c.caps['E_in'] or raw.num_nodes > self.c.caps['N'] or raw.max_nodes > self.max_n or \
                (raw.max_loop_edges if self.flags['self_loop'] else raw.max_in_edges) > self.max_e:
            raise ValueError('batch exceeds the engine capacity')
        self.in_src[:e].copy_(raw.src, non_blocking=True)
        self.in_dst[:e].copy_(raw.dst, non_blocking=True)
        self.in_eptr[:g + 1].copy_(raw.edge_ptr, non_blocking=True)
        self.in_nptr[:g + 1].copy_(raw.node_ptr, non_blocking=True)
        n = raw.num_nodes
        if g < self.G:        # the last, partial batch of an epoch (the reference trains on it): empty trailing graphs; every
            self.in_eptr[g + 1:].fill_(e)     # kernel reads the actual node / edge / graph counts from the device (make_dims)
            self.in_nptr[g + 1:].fill_(n)
        t = self.raw_feat if self.pipeline else self.live   # pipelined: the encoder branch forwards them into the staging set
        t.in_x[:n].copy_(raw.x, non_blocking=True)
        if self.ea_via_raw:                          # the encoder applies E1 to the bond columns (loop rows = 1)
            self.in_ea_raw[:raw.edge_attr.size(0)].copy_(raw.edge_attr, non_blocking=True)
        elif t.in_ea is not None:
            t.in_ea[:raw.edge_attr.size(0)].copy_(raw.edge_attr, non_blocking=True)
        t.in_y[:raw.y.numel()].copy_(raw.y.view(-1), non_blocking=True)

    @torch.no_grad()
    def prime(self, raw):
        """Pipelined mode: encode `raw` into the staging set without training (the first batch of a run)."""
        self.load(raw)
        self._forward_features(self.stage)
        self._encode_and_index(self.stage)
        self._primed = True

    @torch.no_grad()
    def drain(self):
        """Pipelined mode: train on the batch still waiting in the staging set (the last batch of a run)."""
        if not (self.pipeline and self._primed):
            return None
        self.opt.sync_hyper(self._world())
        self.live.buf.copy_(self.stage.buf)
        self._run_train()
        if self.exchange == 'nccl':
            self.opt.all_reduce_grads()
        self._run_opt()
        self._primed = False
        self.steps += 1
        return self.loss

    def _world(self):
        if self.distributed:
            import torch.distributed as dist
            return dist.get_world_size()
        return 1

    def step(self, raw):
        """One full step on `raw`; returns the loss as a device tensor (read it with .item() when needed).
        Pipelined engines encode `raw` while training on the batch handed to the PREVIOUS call and return that batch's
        loss (None for the first call, which only encodes); drain() trains the last one."""
        if self.pipeline and not self._primed:
            self.prime(raw)
            return None
        world = 1
        if self.distributed:
            import torch.distributed as dist
            world = dist.get_world_size()
        self.opt.sync_hyper(world)
        self.load(raw)
        if not self.use_graph or self.steps < 2:     # eager warm-up before the capture
            self._run()
        else:
            if self.graph is None:
                torch.cuda.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._run_main()
                    if self.exchange != 'nccl':
                        self._run_opt()
                if self.exchange == 'nccl':      # the NCCL exchange stays outside the captured graphs
                    self.graph_opt = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph_opt):
                        self._run_opt()
            self.graph.replay()
            if self.exchange == 'nccl':
                self.opt.all_reduce_grads()
                self.graph_opt.replay()
        self.steps += 1
        return self.loss

    def step_read(self, raw):
        """step(raw) with the loss read on the host WITHOUT draining the device: the 4-byte result of this call is copied to a pinned
        slot behind the step's graph (an event marks it), and what is returned is the float of the PREVIOUS call, whose event has
        long fired -- the host stays one launch ahead of the device instead of idling it once per step on `.item()`.
        Returns None until a loss is available (pipelined engines train on the batch of the call before, so the first two calls);
        last_read() hands out the final one."""
        if self._loss_host is None:
            self._loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
            self._loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
            self._loss_calls, self._loss_valid = 0, [False, False]
        out = self.last_read()
        loss = self.step(raw)
        k = self._loss_calls & 1
        self._loss_valid[k] = loss is not None
        if loss is not None:
            self._loss_host[k:k + 1].copy_(loss.view(-1)[:1], non_blocking=True)
            self._loss_ev[k].record(torch.cuda.current_stream(self.c.dev))
        self._loss_calls += 1
        return out

    def last_read(self):
        """The loss of the latest step_read() call as a float (waits for that step only); None if it produced none."""
        if self._loss_host is None or self._loss_calls == 0:
            return None
        k = (self._loss_calls - 1) & 1
        if not self._loss_valid[k]:
            return None
        self._loss_ev[k].synchronize()
        return float(self._loss_host[k])

    def profile(self, raw, reps=10, flush=None):
        """Per-kernel DEVICE times of the captured step: the same launch sequence captured once more on a single stream
        with a graph-capturable event after every launch, replayed `reps` times (parameters are restored afterwards).
        Returns ({label: ms per step}, {label: launches per step})."""
        import copy
        saved = (self.opt.flat.clone(), self.opt.exp_avg.clone(), self.opt.exp_avg_sq.clone(), self.opt.state.clone(),
                 [(b.running_mean.clone(), b.running_var.clone()) for b in self._bns])
        self.load(raw)
        torch.cuda.synchronize()
        _lib.PROFILE, _lib.PROFILE_EXTERNAL, self.inline_branches = [], True, True
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._run_main()
            self.opt.step_device()               # rank-local: no exchange inside the profiling replay
            _lib.mark('end')
        marks, _lib.PROFILE, _lib.PROFILE_EXTERNAL, self.inline_branches = _lib.PROFILE, None, False, False
        ms, calls = {}, {}
        for _ in range(reps):
            if flush is not None:
                flush.zero_()
            g.replay()
            torch.cuda.synchronize()
            for (l0, e0), (l1, e1) in zip(marks[:-1], marks[1:]):
                ms[l1] = ms.get(l1, 0.0) + e0.elapsed_time(e1) / reps
                calls[l1] = calls.get(l1, 0) + 1.0 / reps
        self.opt.flat.copy_(saved[0]); self.opt.exp_avg.copy_(saved[1]); self.opt.exp_avg_sq.copy_(saved[2])
        self.opt.state.copy_(saved[3])
        for b, (rm, rv) in zip(self._bns, saved[4]):
            b.running_mean.copy_(rm); b.running_var.copy_(rv)
        return ms, calls

    def sync_counters(self):
        """BatchNorm `num_batches_tracked` is bookkeeping only (momentum is fixed): advanced lazily, outside the graph."""
        d = self.steps - self._bn_synced
        if d:
            for bn in self._bns:
                bn.num_batches_tracked.add_(d)
            self._bn_synced = self.steps

    def check_errors(self):
        """Lazy data-error check (degree >= 200, bad ids, capacity) over EVERY batch since the last call: one sync, call it
        once per epoch.  The encoder's per-call counters are zeroed each step, the sticky slots are not."""
        cnt = self.counters.cpu()
        self.counters[_lib.CTR_STICKY_ERROR:_lib.CTR_MAX_NNZ + 1].zero_()
        _lib.raise_data_errors(int(cnt[_lib.CTR_STICKY_ERROR]) | int(cnt[1]))
        worst = max(int(cnt[_lib.CTR_MAX_NNZ]), int(cnt[0]))
        if worst > self.rec.numel():
            raise RuntimeError('engine record capacity exceeded: %d > %d (the edges without room were trained on as empty bags; '
                               'raise records_per_edge)' % (worst, self.rec.numel()))
        if int(self.idx_err.cpu()[0]):
            raise RuntimeError('edge_index out of range after collation')
        if self.opt.peers is not None and self.opt.peers.timed_out():
            raise RuntimeError('gradient exchange: a peer did not arrive within 4 s (ranks must run the same number of steps)')
