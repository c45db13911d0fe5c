import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gemm_gpu import run_gemm
g = torch.Generator(device='cuda').manual_seed(0)
for (M, N, K) in [(128, 32, 32), (128, 32, 64), (128, 64, 32), (128, 128, 32), (256, 128, 512), (128, 32, 512), (256, 256, 256), (1000, 288, 256), (300, 160, 96), (4096, 64, 40)]:
    for a_mn, b_mn in ((False, False), (False, True), (True, False), (True, True)):
        if (a_mn and M % 4) or (b_mn and N % 4):
            continue
        A = torch.randn(M, K, device='cuda', generator=g); B = torch.randn(N, K, device='cuda', generator=g)
        ref = A.double() @ B.double().t()
        C = run_gemm(A, B, a_mn, b_mn)
        torch.cuda.synchronize()
        rel = ((C.double() - ref).norm() / ref.norm()).item()
        print(M, N, K, 'a_mn' if a_mn else 'a_k ', 'b_mn' if b_mn else 'b_k ', '%.2e' % rel, 'nan' if not torch.isfinite(C).all() else '')
