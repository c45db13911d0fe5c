"""Flat-buffer Adam on the sm_100a `adam_step` kernel, plus the data-parallel gradient exchange.

All parameters are re-pointed at views of ONE contiguous fp32 buffer (and their .grad at views of one gradient
buffer), so a step is one kernel launch and the data-parallel exchange is one NCCL all-reduce over NVLink.
Update rule = torch.optim.Adam defaults (reference: run_graphcount.py:478, run_zinc.py:263, run_ogb_mol.py:436).
"""
import ctypes

import torch

from . import _lib


class _RawCuda(object):
    """CUDA array interface over a raw device allocation of the library (zero-copy torch view of peer-shareable memory)."""
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = dict(shape=(n, ), typestr=typestr, data=(int(ptr), False), version=2)


class PeerBuffers(object):
    """Gradient / parameter / flag buffers of every rank of the node, opened through CUDA IPC (one process per GPU), for the
    fused exchange kernel escgnn_allreduce_adam (csrc/p2p.cu).  Collective: every rank of `group` must construct it."""

    def __init__(self, n_floats, device, group=None):
        import torch.distributed as dist
        L = _lib.lib()
        self.L, self.world, self.rank = L, dist.get_world_size(group), dist.get_rank(group)
        self.n = int(n_floats)
        flag_words = int(L.escgnn_p2p_flag_words(self.world))
        sizes = dict(grad=4 * self.n, param=4 * self.n, flags=8 * flag_words)
        self.own, handles = {}, {}
        for k, b in sizes.items():
            ptr, h = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
            _lib.check(L.escgnn_p2p_alloc(b, ctypes.byref(ptr), h), 'p2p_alloc')
            self.own[k], handles[k] = ptr.value, bytes(h)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=group)
        self.peer, self._opened = {k: [0] * self.world for k in sizes}, []
        failed = 0
        for r, hs in enumerate(gathered):
            for k in sizes:
                if r == self.rank:
                    self.peer[k][r] = self.own[k]
                else:
                    ptr = ctypes.c_void_p()
                    rc = L.escgnn_p2p_open((ctypes.c_ubyte * 64).from_buffer_copy(hs[k]), ctypes.byref(ptr))
                    if rc != 0:
                        failed = rc
                        continue
                    self.peer[k][r] = ptr.value
                    self._opened.append(ptr.value)
        # every rank takes the same decision: one rank that cannot map a peer (IPC disabled, no peer access) fails all of them
        flag = torch.tensor([1 if failed else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag[0]):
            self.close()
            for ptr in self.own.values():
                L.escgnn_p2p_free(ctypes.c_void_p(ptr))
            raise RuntimeError('peer-memory exchange unavailable on this node (cudaIpcOpenMemHandle failed on some rank: %d)' % failed)
        self.arrays = {k: (ctypes.c_void_p * self.world)(*v) for k, v in self.peer.items()}     # host arrays of device pointers
        self.grad = torch.as_tensor(_RawCuda(self.own['grad'], self.n, '<f4'), device=device)
        self.param = torch.as_tensor(_RawCuda(self.own['param'], self.n, '<f4'), device=device)
        self.flags = torch.as_tensor(_RawCuda(self.own['flags'], flag_words, '<i8'), device=device)
        dist.barrier(group=group)

    def timed_out(self):
        block = 2 * self.world + 8
        return any(int(self.flags[b * block + 2 * self.world + 2].item()) != 0 for b in range(4))

    def close(self):
        for p in self._opened:
            self.L.escgnn_p2p_close(ctypes.c_void_p(p))
        self._opened = []


class FlatAdam(object):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, first=(), p2p_group=False, last=()):
        """`first`: parameters to lay out first and in this order (tensors the engine wants adjacent in memory, e.g. the
        edge-projection weights of all layers so they form ONE [sum C_in, edge_dim] operand); the rest keep their order.
        `p2p_group`: None / a process group = keep parameters and gradients in peer-shareable memory and run the data-parallel
        exchange fused with the update (`step_exchange_device`); False = local buffers (+ `all_reduce_grads` over NCCL)."""
        params = [p for p in params if p.requires_grad]
        head = [p for p in first if p.requires_grad]
        tail = [p for p in last if p.requires_grad]             # `last`: laid out at the END (the engine puts the parameters whose
        ids = {id(p) for p in head} | {id(p) for p in tail}     # gradients arrive last there: one contiguous late exchange bucket)
        self.params = head + [p for p in params if id(p) not in ids] + tail
        if not self.params or not self.params[0].is_cuda:
            raise RuntimeError('FlatAdam needs CUDA parameters; there is no CPU fallback')
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        pad = [(-s) % 4 for s in sizes]                       # keep every tensor 16-byte aligned in the flat buffer
        total = sum(s + q for s, q in zip(sizes, pad))
        self.peers = None
        if p2p_group is not False:
            self.peers = PeerBuffers(total, dev, group=p2p_group)
            self.flat, self.grad = self.peers.param, self.peers.grad
        else:
            self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
            self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        off = 0
        self.tail_begin = total - sum(s + q for (p, s, q) in zip(self.params, sizes, pad) if any(p is t for t in tail))
        for p, s, q in zip(self.params, sizes, pad):
            self.flat[off:off + s].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + s].view(p.shape)
            p.grad = self.grad[off:off + s].view(p.shape)
            off += s + q
        self.lr, self.betas, self.eps, self.t = lr, betas, eps, 0
        self.param_groups = [dict(lr=lr)]
        # device-side hyper-parameters / step counter (graph-capturable path)
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, 1.0, 1.0, 1.0, 0.0], dtype=torch.float32, device=dev)
        self.state = torch.zeros(1, dtype=torch.int64, device=dev)
        self._hyper_host = (lr, 1.0)

    def zero_grad(self, set_to_none=False):
        self.grad.zero_()
        for p in self.params:                                 # autograd may have replaced .grad; keep the views
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr() or \
                    p.grad.data_ptr() >= self.grad.data_ptr() + self.grad.numel() * 4:
                off = (p.data.data_ptr() - self.flat.data_ptr()) // 4
                p.grad = self.grad[off:off + p.numel()].view(p.shape)

    def all_reduce_grads(self, group=None):
        """Data-parallel exchange: one sum all-reduce of the flat gradient (NCCL over NVLink); step() rescales."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1

    def step_device(self, world_size=1):
        """Adam with the step counter in device memory: identical launches every step, safe inside a CUDA graph.
        (A learning-rate or world-size change is pushed with one small copy OUTSIDE the captured region.)"""
        st = ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().escgnn_adam_step_device(P(self.flat), P(self.grad), P(self.exp_avg), P(self.exp_avg_sq),
                                                      self.flat.numel(), P(self.hyper), P(self.state), st),
                   'adam_step_device')

    def step_exchange_device(self, begin=0, end=None, bucket=0, tick=True):
        """Gradient exchange over NVLink peer memory + Adam in one graph-capturable launch (csrc/p2p.cu) for the element range
        [begin, end): every rank must issue the same sequence of calls.  Averages the gradients (hyper[4] = 1 / world, see
        sync_hyper).  A step may be split into buckets (`bucket` 0..3 = independent barrier flags): `tick=True` on the first call
        of the step only, and the later calls must be ordered after it on the device."""
        pb = self.peers
        end = self.flat.numel() if end is None else end
        st = ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().escgnn_allreduce_adam_range(pb.arrays['grad'], pb.arrays['param'], pb.arrays['flags'], pb.rank, pb.world,
                                                          int(begin), int(end), int(bucket), int(bool(tick)), P(self.exp_avg),
                                                          P(self.exp_avg_sq), P(self.hyper), P(self.state), st), 'allreduce_adam')

    def sync_hyper(self, world_size=1):
        want = (self.param_groups[0]['lr'], 1.0 / world_size)
        if want != self._hyper_host:
            self.hyper[0] = want[0]
            self.hyper[4] = want[1]
            self._hyper_host = want

    def step(self, world_size=1):
        self.t += 1
        lr = self.param_groups[0]['lr']
        st = ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().escgnn_adam_step(P(self.flat), P(self.grad), P(self.exp_avg), P(self.exp_avg_sq),
                                               self.flat.numel(), lr, self.betas[0], self.betas[1], self.eps, self.t,
                                               1.0 / world_size, st), 'adam_step')
