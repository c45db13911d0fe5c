"""128 x 256 tiles (escgnn_gemm_set_wide) against 128 x 128 tiles on the edge-level products of a reference batch: us per launch
(CUDA graph of 30 back-to-back launches, L2-warm) and useful TFLOP/s."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esc_gnn_b200 import _lib
L = _lib.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ws = torch.empty(8 << 20, device='cuda')


def run(A, B, b_mn, M, N, K, reps=30):
    C = torch.empty(M, N, device='cuda')
    f = lambda: _lib.check(L.escgnn_gemm_tf32x3(P(A), A.stride(0), 0, P(B), B.stride(0), b_mn, P(C), N, None, M, N, K, 0, P(ws), ws.numel(), st()), 'g')
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


for M in (12092, 12800, 8320, 18900):
    for name, K, b_mn in (('fwd   z_embedding', 256, 0), ('dgrad z_embedding', 256, 1), ('dgrad projection ', 1056, 1), ('fwd   K=288      ', 288, 0)):
        A = torch.randn(M, K, device='cuda')
        B = torch.randn(K, 256, device='cuda') if b_mn else torch.randn(256, K, device='cuda')
        t = {}
        for wide in (0, 1):
            L.escgnn_gemm_set_wide(wide)
            t[wide] = run(A, B, b_mn, M, 256, K)
        L.escgnn_gemm_set_wide(1)
        fl = 2.0 * M * 256 * K
        print('%s M=%6d N= 256 K=%5d   128-wide %6.1f us (%5.1f TF)   256-wide %6.1f us (%5.1f TF)' % (
            name, M, K, t[0], fl / t[0] / 1e6, t[1], fl / t[1] / 1e6), flush=True)
