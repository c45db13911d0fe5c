"""For ncu: plain GEMM, fused Linear+BN+act forward and backward on the N-level shape, plain GEMM on the E-level shape."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esc_gnn_b200 import _lib
L = _lib.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def go(rows_cap, rows, n, k, fused):
    x = torch.randn(rows_cap, k, device='cuda'); x[rows:] = 0
    w = torch.randn(n, k, device='cuda') / k ** 0.5
    v = [torch.ones(n, device='cuda') for _ in range(8)]
    y, out = torch.empty(rows_cap, n, device='cuda'), torch.empty(rows_cap, n, device='cuda')
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    ws = torch.zeros(L.escgnn_linear_bn_workspace_floats(rows_cap, n), device='cuda')
    gws = torch.empty(8 << 20, device='cuda')
    for _ in range(2):
        _lib.check(L.escgnn_gemm_tf32x3_bounded(P(x), k, 0, P(w), k, 0, P(y), n, P(v[0]), rows_cap, n, k, 0, P(gws), gws.numel(), P(d_rows), 1, st()), 'g')
        if fused:
            _lib.check(L.escgnn_linear_bn_act_fwd(P(x), k, P(w), k, P(v[0]), rows_cap, n, k, P(d_rows), P(v[1]), P(v[2]), P(v[3]), P(v[4]), P(v[5]), P(v[6]),
                                                  2, 1e-5, 0.1, P(y), n, P(out), n, P(ws), ws.numel(), st()), 'f')
            _lib.check(L.escgnn_linear_bn_act_bwd(P(x), k, P(w), n, rows_cap, n, k, P(d_rows), P(y), n, P(v[5]), P(v[6]), P(v[1]), P(v[2]), 2, n, P(v[3]),
                                                  P(v[4]), P(out), n, P(ws), ws.numel(), st()), 'f')
    torch.cuda.synchronize()
go(6302, 5906, 256, 256, True)
go(12847, 12092, 256, 256, False)
print('ok')
