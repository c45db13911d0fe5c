"""A few engine-shaped GEMMs for ncu: fwd, dgrad, wgrad of an E x 288 -> 256 Linear (E = 12800), then fwd / dgrad of the node-level
256 -> 256 Linear of a reference batch (5906 rows: the default one-CTA-per-SM instantiation <128,0,B_MN,4,4,0,8,2>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gemm_gpu import run_gemm
g = torch.Generator(device='cuda').manual_seed(0)
E = 12800
X = torch.randn(E, 288, device='cuda', generator=g); W = torch.randn(256, 288, device='cuda', generator=g)
dY = torch.randn(E, 256, device='cuda', generator=g)
for _ in range(3):
    Y = run_gemm(X, W, False, False)                          # fwd   [E,256] = X W^T
    dX = run_gemm(dY, W.t().contiguous(), False, True)        # dgrad [E,288] = dY W      (B stored [K=256, N=288])
    dW = run_gemm(dY.t().contiguous(), X.t().contiguous(), True, True)   # wgrad [256,288] = dY^T X (both stored [rows, C])
Xn = torch.randn(5906, 256, device='cuda', generator=g); Wn = torch.randn(256, 256, device='cuda', generator=g)
for _ in range(3):
    Yn = run_gemm(Xn, Wn, False, False)
    dXn = run_gemm(Xn, Wn.t().contiguous(), False, True)
torch.cuda.synchronize()
print('ok', float(Y.abs().mean()), float(dX.abs().mean()), float(dW.abs().mean()))
