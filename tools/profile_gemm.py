"""A few engine-shaped GEMMs for ncu (fwd, dgrad, wgrad of an E x 288 -> 256 Linear)."""
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
torch.cuda.synchronize()
print('ok', float(Y.abs().mean()), float(dX.abs().mean()), float(dW.abs().mean()))
