"""Back-to-back timing of engine-shaped GEMMs (CUDA events, L2-warm): us per launch and useful TFLOP/s."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esc_gnn_b200 import _lib
L = _lib.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ws = torch.empty(8 << 20, device='cuda')
def run(name, A, a_mn, B, b_mn, M, N, K, reps=30):
    C = torch.empty(M, N, device='cuda')
    f = lambda: _lib.check(L.escgnn_gemm_tf32x3(P(A), A.stride(0), a_mn, P(B), B.stride(0), b_mn, P(C), N, None, M, N, K, 0, P(ws), ws.numel(), st()), 'g')
    for _ in range(3): f()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    torch.backends.cuda.matmul.allow_tf32 = False
    print('%-34s M=%6d N=%4d K=%6d  %7.1f us  %6.1f useful TFLOP/s' % (name, M, N, K, us, 2.0 * M * N * K / us / 1e6))
E, Nn = 12800, 6300
plan = int(sys.argv[1]) if len(sys.argv) > 1 else -1
L.escgnn_gemm_set_plan(plan)
print('plan', plan)
for rows, tag in ((E, 'E'), (Nn, 'N')):
    for cin, cout in ((288, 256), (256, 256)):
        X = torch.randn(rows, cin, device='cuda'); W = torch.randn(cout, cin, device='cuda'); dY = torch.randn(rows, cout, device='cuda')
        run('fwd   %s x %d -> %d' % (tag, cin, cout), X, 0, W, 0, rows, cout, cin)
        run('dgrad %s x %d -> %d' % (tag, cout, cin), dY, 0, W, 1, rows, cin, cout)
        run('wgrad %s rows, %d x %d' % (tag, cout, cin), dY, 1, X, 1, cout, cin, rows)
# grouped edge-projection shapes of the engine (all layers' conv.lin in one product)
Z = torch.randn(E, 288, device='cuda'); Wc = torch.randn(800, 288, device='cuda'); dEE = torch.randn(E, 800, device='cuda')
run('grouped fwd   E x 288 -> 800', Z, 0, Wc, 0, E, 800, 288)
run('grouped dgrad E x 800 -> 288', dEE, 0, Wc, 1, E, 288, 800)
run('grouped wgrad E rows, 800 x 288', dEE, 1, Z, 1, 800, 288, E)
def run_atomic(name, A, B, M, N, K, reps=30):
    C = torch.zeros(M, N, device='cuda')
    f = lambda: _lib.check(L.escgnn_gemm_tf32x3(P(A), A.stride(0), 1, P(B), B.stride(0), 1, P(C), N, None, M, N, K, 2, None, 0, st()), 'g')
    for _ in range(3): f()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    print('%-34s M=%6d N=%4d K=%6d  %7.1f us  %6.1f useful TFLOP/s  (atomic split-K)' % (name, M, N, K, us, 2.0 * M * N * K / us / 1e6))
run_atomic('grouped wgrad E rows, 800 x 288', dEE, Z, 800, 288, E)
dYn = torch.randn(Nn, 256, device='cuda'); Xn = torch.randn(Nn, 256, device='cuda')
run_atomic('wgrad N rows, 256 x 256', dYn, Xn, 256, 256, Nn)
