"""One process for ncu: engine-shaped GEMMs (fwd / dgrad / wgrad) followed by the one-launch BatchNorm kernels (N- and E-level)."""
import os, sys, ctypes, runpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profile_gemm.py'))
import torch
from esc_gnn_b200 import _lib
L = _lib.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for rows, cap in ((5906, 6302), (12092, 12847)):
    C = 256
    x = torch.randn(cap, C, device='cuda'); y = torch.empty_like(x); dy = torch.randn_like(x); dx = torch.empty_like(x)
    g = torch.ones(C, device='cuda'); b = torch.zeros(C, device='cuda'); rm = torch.zeros(C, device='cuda'); rv = torch.ones(C, device='cuda')
    mean = torch.zeros(C, device='cuda'); rstd = torch.ones(C, device='cuda'); dg = torch.zeros(C, device='cuda'); db = torch.zeros(C, device='cuda')
    part = torch.zeros(L.escgnn_dense_partial_floats(cap, C), device='cuda')
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    for _ in range(2):
        _lib.check(L.escgnn_bn_act_fwd(P(x), C, P(g), P(b), P(rm), P(rv), P(mean), P(rstd), P(part), 2, 1e-5, 0.1, 1, P(d_rows), cap, C, P(y), C, st()), 'bn_act_fwd')
        _lib.check(L.escgnn_bn_act_bwd(P(x), C, P(dy), C, None, 0, P(mean), P(rstd), P(g), P(b), 2, 1, P(part), P(d_rows), cap, C, P(dg), P(db), P(dx), C, st()), 'bn_act_bwd')
torch.cuda.synchronize()
print('bn ok', float(y.abs().mean()), float(dx.abs().mean()))
