"""Fused Linear+BatchNorm+act launch vs GEMM + stand-alone BatchNorm on the engine's shapes (CUDA events, back-to-back in a graph,
L2-warm), and the accuracy of the tcgen05 3xTF32 product against fp64 (mean SIGNED relative error = accumulation bias, rms).

    python tools/bench_linear_bn.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from esc_gnn_b200 import _lib

L = _lib.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.backends.cuda.matmul.allow_tf32 = False


def timeit(f, reps=30):
    for _ in range(3):
        f()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def bench(rows_cap, rows, n_out, k_in):
    x = torch.randn(rows_cap, k_in, device='cuda'); x[rows:] = 0
    w = torch.randn(n_out, k_in, device='cuda') / k_in ** 0.5
    bias = torch.randn(n_out, device='cuda')
    gamma, beta = torch.ones(n_out, device='cuda'), torch.zeros(n_out, device='cuda')
    rm, rv = torch.zeros(n_out, device='cuda'), torch.ones(n_out, device='cuda')
    mean, rstd = torch.zeros(n_out, device='cuda'), torch.zeros(n_out, device='cuda')
    y, out = torch.empty(rows_cap, n_out, device='cuda'), torch.empty(rows_cap, n_out, device='cuda')
    dy = torch.randn(rows_cap, n_out, device='cuda'); dy[rows:] = 0
    dx = torch.empty(rows_cap, k_in, device='cuda')
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    ws = torch.zeros(L.escgnn_linear_bn_workspace_floats(rows_cap, max(n_out, k_in)), device='cuda')
    part = torch.zeros(L.escgnn_dense_partial_floats(rows_cap, 2048), device='cuda')
    gws = torch.empty(8 << 20, device='cuda')
    gemm = lambda: _lib.check(L.escgnn_gemm_tf32x3_bounded(P(x), k_in, 0, P(w), k_in, 0, P(y), n_out, P(bias), rows_cap, n_out, k_in, 0, P(gws),
                                                           gws.numel(), P(d_rows), 1, st()), 'g')
    bn = lambda: _lib.check(L.escgnn_bn_act_fwd(P(y), n_out, P(gamma), P(beta), P(rm), P(rv), P(mean), P(rstd), P(part), 2, 1e-5, 0.1, 1,
                                                P(d_rows), rows_cap, n_out, P(out), n_out, st()), 'b')
    fused = lambda yy: _lib.check(L.escgnn_linear_bn_act_fwd(P(x), k_in, P(w), k_in, P(bias), rows_cap, n_out, k_in, P(d_rows), P(gamma), P(beta),
                                                             P(rm), P(rv), P(mean), P(rstd), 2, 1e-5, 0.1, P(yy), n_out, P(out), n_out, P(ws),
                                                             ws.numel(), st()), 'f')
    dgrad = lambda: _lib.check(L.escgnn_gemm_tf32x3_bounded(P(dy), n_out, 0, P(w), k_in, 1, P(dx), k_in, None, rows_cap, k_in, n_out, 0, P(gws),
                                                            gws.numel(), P(d_rows), 1, st()), 'g')
    xs = torch.randn(rows_cap, k_in, device='cuda')
    dxo = torch.empty(rows_cap, k_in, device='cuda')
    m2, r2 = torch.zeros(k_in, device='cuda'), torch.ones(k_in, device='cuda')
    g2, b2 = torch.ones(k_in, device='cuda'), torch.zeros(k_in, device='cuda')
    dg, db = torch.zeros(k_in, device='cuda'), torch.zeros(k_in, device='cuda')
    bnb = lambda: _lib.check(L.escgnn_bn_act_bwd(P(xs), k_in, P(dx), k_in, None, 0, P(m2), P(r2), P(g2), P(b2), 2, 1, P(part), P(d_rows), rows_cap,
                                                 k_in, P(dg), P(db), P(dxo), k_in, st()), 'b')
    fusedb = lambda: _lib.check(L.escgnn_linear_bn_act_bwd(P(dy), n_out, P(w), k_in, rows_cap, k_in, n_out, P(d_rows), P(xs), k_in, P(m2), P(r2),
                                                           P(g2), P(b2), 2, k_in, P(dg), P(db), P(dxo), k_in, P(ws), ws.numel(), st()), 'f')
    ok = L.escgnn_linear_bn_fusable(rows_cap, n_out, k_in)
    t_g, t_b = timeit(gemm), timeit(bn)
    t_pair = timeit(lambda: (gemm(), bn()))
    line = 'rows %6d/%6d  %4d -> %4d | gemm %5.1f  bn %5.1f  pair %5.1f' % (rows, rows_cap, k_in, n_out, t_g, t_b, t_pair)
    if ok:
        line += ' | fused %5.1f  fused(no y) %5.1f' % (timeit(lambda: fused(y)), timeit(lambda: fused(None)))
    t_d, t_bb = timeit(dgrad), timeit(bnb)
    line += ' || dgrad %5.1f  bn_bwd %5.1f  pair %5.1f' % (t_d, t_bb, timeit(lambda: (dgrad(), bnb())))
    if L.escgnn_linear_bn_fusable(rows_cap, k_in, n_out):
        line += ' | fused %5.1f' % timeit(fusedb)
    print(line + '  (us)')


def accuracy(M, N, K):
    g = torch.Generator(device='cuda').manual_seed(1)
    A = torch.randn(M, K, device='cuda', generator=g).abs() + 0.1        # positive operands: no cancellation, error relative to the result
    B = torch.randn(N, K, device='cuda', generator=g).abs() + 0.1
    ref = A.double() @ B.double().t()
    C = torch.empty(M, N, device='cuda')
    gws = torch.empty(8 << 20, device='cuda')
    _lib.check(L.escgnn_gemm_tf32x3(P(A), K, 0, P(B), K, 0, P(C), N, None, M, N, K, 0, P(gws), gws.numel(), st()), 'g')
    T = A @ B.t()
    L.escgnn_gemm_set_drain(0)
    C0 = torch.empty(M, N, device='cuda')
    _lib.check(L.escgnn_gemm_tf32x3(P(A), K, 0, P(B), K, 0, P(C0), N, None, M, N, K, 0, P(gws), gws.numel(), st()), 'g')
    L.escgnn_gemm_set_drain(2)
    for name, X in (('tcgen05 3xTF32 drain', C), ('tcgen05 in-TC accum', C0), ('cuBLAS fp32', T)):
        rel = (X.double() - ref) / ref
        print('  %-22s M %5d N %4d K %5d: mean signed rel err %+.3e   rms %.3e   max %.3e' % (name, M, N, K, rel.mean().item(), rel.pow(2).mean().sqrt().item(),
                                                                                              rel.abs().max().item()))


def speed():
    gws = torch.empty(8 << 20, device='cuda')
    for M, N, K in ((6302, 256, 256), (12847, 256, 256), (12847, 288, 1056), (12847, 1056, 288)):
        A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda'); C = torch.empty(M, N, device='cuda')
        line = 'gemm M %6d N %5d K %5d:' % (M, N, K)
        for mode in (0, 1, 2):
            L.escgnn_gemm_set_drain(mode)
            f = lambda: _lib.check(L.escgnn_gemm_tf32x3(P(A), K, 0, P(B), K, 0, P(C), N, None, M, N, K, 0, P(gws), gws.numel(), st()), 'g')
            line += '  drain=%d %6.1f us' % (mode, timeit(f))
        L.escgnn_gemm_set_drain(2)
        print(line)


if __name__ == '__main__':
    speed()
    for shape in ((6302, 5906, 256, 256), (6302, 5906, 256, 32), (6302, 128, 256, 256), (12847, 12092, 256, 256), (2600, 2500, 256, 256)):
        bench(*shape)
    for shape in ((1024, 256, 64), (1024, 256, 256), (1024, 256, 1024), (256, 256, 8192)):
        accuracy(*shape)
