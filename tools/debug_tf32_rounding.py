import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esc_gnn_b200 import _lib
L = _lib.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device='cuda').manual_seed(0)
M, N, K = 256, 128, 512
A = torch.randn(M, K, device='cuda', generator=g); B = torch.randn(N, K, device='cuda', generator=g)
ref = A.double() @ B.double().t()
def hi(x, mode):
    b = x.view(torch.int32)
    if mode == 'trunc': h = b & -8192
    elif mode == 'rna': h = (b + 0x1000) & -8192
    else: h = (b + 0xfff + ((b >> 13) & 1)) & -8192
    return h.view(torch.float32)
def run(Ah, Al, Bh, Bl):
    C = torch.zeros(M, N, device='cuda')
    _lib.check(L.escgnn_gemm_tf32x3(P(Ah), P(Al), K, 0, P(Bh), P(Bl), K, 0, P(C), N, None, M, N, K, 0, None, 0, st), 'gemm_tf32x3')
    return ((C.double() - ref).norm() / ref.norm()).item()
z = torch.zeros_like
print('single pass (lo = 0)          ', run(A, z(A), B, z(B)))
for mode in ('trunc', 'rna', 'rne'):
    print('raw hi, lo = a - %-5s(a)      ' % mode, run(A, (A - hi(A, mode)).contiguous(), B, (B - hi(B, mode)).contiguous()))
    ah, bh = hi(A, mode).contiguous(), hi(B, mode).contiguous()
    print('explicit hi=%-5s, lo = a - hi  ' % mode, run(ah, (A - ah).contiguous(), bh, (B - bh).contiguous()))
    al, bl = (A - ah).contiguous(), (B - bh).contiguous()
    print('explicit hi=%-5s, lo pre-rounded' % mode, run(ah, hi(al, mode).contiguous(), bh, hi(bl, mode).contiguous()))
torch.backends.cuda.matmul.allow_tf32 = False
print('cuBLAS fp32', (((A @ B.t()).double() - ref).norm() / ref.norm()).item())
