"""GPU: the tcgen05 3xTF32 GEMM (all operand majors, tails, split-K, bias, accumulate) against an fp64 torch product."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _lo(x):
    from esc_gnn_b200 import _lib
    lo = torch.empty_like(x)
    _lib.check(_lib.lib().escgnn_tf32_split_lo(_p(x), x.stride(0), _p(lo), lo.stride(0), x.size(0), x.size(1), _st()), 'tf32_split_lo')
    return lo


def run_gemm(A, B, a_mn, b_mn, bias=None, C0=None, simple=False):
    """A: [M,K] logical, B: [N,K] logical. Storage follows the major flags."""
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    M, K = A.shape
    N = B.shape[0]
    As = A.t().contiguous() if a_mn else A.contiguous()
    Bs = B.t().contiguous() if b_mn else B.contiguous()
    C = C0.clone() if C0 is not None else torch.full((M, N), float('nan'), device='cuda')
    if simple:
        _lib.check(L.escgnn_gemm_simple(_p(As), As.stride(0), int(a_mn), _p(Bs), Bs.stride(0), int(b_mn), _p(C), N, _p(bias),
                                        M, N, K, int(C0 is not None), _st()), 'gemm_simple')
        return C
    ws_n = L.escgnn_gemm_workspace_floats(M, N, K)
    ws = torch.empty(max(ws_n, 1), device='cuda')
    _lib.check(L.escgnn_gemm_tf32x3(_p(As), As.stride(0), int(a_mn), _p(Bs), Bs.stride(0), int(b_mn), _p(C), N, _p(bias), M, N, K,
                                    int(C0 is not None), _p(ws) if ws_n else None, ws_n, _st()), 'gemm_tf32x3')
    return C


def check(A, B, a_mn, b_mn, bias=None, C0=None, simple=False):
    C = run_gemm(A, B, a_mn, b_mn, bias, C0, simple)
    ref = A.double() @ B.double().t()
    if bias is not None:
        ref = ref + bias.double()
    if C0 is not None:
        ref = ref + C0.double()
    scale = (A.double().abs() @ B.double().abs().t()).max().item()
    err = (C.double() - ref).abs().max().item()
    assert torch.isfinite(C).all()
    assert err <= 4e-6 * scale, (err, scale, err / scale)


@pytest.mark.parametrize('a_mn,b_mn', [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize('M,N,K', [(128, 32, 32), (256, 256, 256), (1000, 288, 256), (300, 160, 96), (4096, 64, 40)])
def test_gemm_tf32x3_matches_fp64(a_mn, b_mn, M, N, K):
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    A = torch.randn(M, K, device='cuda', generator=g)
    B = torch.randn(N, K, device='cuda', generator=g)
    if (a_mn and M % 4) or (b_mn and N % 4) or (not a_mn and K % 4) or (not b_mn and K % 4):
        pytest.skip('pitch not 16-byte aligned for this major')
    check(A, B, a_mn, b_mn)


def test_gemm_bias_accumulate_and_splitk():
    g = torch.Generator(device='cuda').manual_seed(1)
    A = torch.randn(12800, 256, device='cuda', generator=g)      # wgrad shape: dW[256,288] = dY^T X over 12.8k rows
    X = torch.randn(12800, 288, device='cuda', generator=g)
    check(A.t().contiguous(), X.t().contiguous(), True, True)     # logical [256,12800] x [288,12800]^T, stored row-major [rows, C]
    Aw = torch.randn(500, 256, device='cuda', generator=g)
    W = torch.randn(288, 256, device='cuda', generator=g)
    bias = torch.randn(288, device='cuda', generator=g)
    check(Aw, W, False, False, bias=bias)
    check(Aw, W, False, False, bias=bias, C0=torch.randn(500, 288, device='cuda', generator=g))
    check(Aw, W.t().contiguous().t(), False, False)


def test_gemm_simple_fallback_any_shape():
    g = torch.Generator(device='cuda').manual_seed(2)
    for (M, N, K, a_mn, b_mn) in [(333, 10, 256, False, False), (333, 256, 10, False, True), (10, 256, 333, True, True), (77, 1, 33, False, False)]:
        A = torch.randn(M, K, device='cuda', generator=g)
        B = torch.randn(N, K, device='cuda', generator=g)
        C = run_gemm(A, B, a_mn, b_mn, simple=True)
        torch.testing.assert_close(C, A @ B.t(), rtol=1e-4, atol=1e-4)


def test_gemm_precision_beats_single_pass_tf32():
    g = torch.Generator(device='cuda').manual_seed(3)
    A = torch.randn(512, 1024, device='cuda', generator=g)
    B = torch.randn(256, 1024, device='cuda', generator=g)
    C = run_gemm(A, B, False, False)
    ref = A.double() @ B.double().t()
    rel = ((C.double() - ref).norm() / ref.norm()).item()
    torch.backends.cuda.matmul.allow_tf32 = True
    rel_tf32 = (((A @ B.t()).double() - ref).norm() / ref.norm()).item()
    torch.backends.cuda.matmul.allow_tf32 = False
    rel_fp32 = (((A @ B.t()).double() - ref).norm() / ref.norm()).item()
    assert rel < 5e-6 and rel < rel_tf32 / 20, (rel, rel_tf32, rel_fp32)


@pytest.mark.parametrize('M,N,K,a_mn,b_mn', [(256, 288, 12800, True, True), (800, 288, 6000, True, True), (256, 256, 5000, False, False),
                                             (128, 40, 4096, True, True)])
def test_gemm_atomic_split_k_accumulates_in_place(M, N, K, a_mn, b_mn):
    """accumulate = 2: the K-slices of a weight-gradient-shaped product are added into C with vector reductions (no
    workspace, no reduction launch); C keeps its initial value plus the product, bias added exactly once."""
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    A = torch.randn(M, K, device='cuda', generator=g); B = torch.randn(N, K, device='cuda', generator=g)
    bias = torch.randn(N, device='cuda', generator=g)
    C0 = torch.randn(M, N, device='cuda', generator=g)
    As = A.t().contiguous() if a_mn else A.contiguous()
    Bs = B.t().contiguous() if b_mn else B.contiguous()
    C = C0.clone()
    _lib.check(L.escgnn_gemm_tf32x3(_p(As), As.stride(0), int(a_mn), _p(Bs), Bs.stride(0), int(b_mn), _p(C), N, _p(bias), M, N, K, 2,
                                    None, 0, _st()), 'gemm_tf32x3')
    ref = A.double() @ B.double().t() + bias.double() + C0.double()
    scale = (A.double().abs() @ B.double().abs().t()).max().item()
    assert (C.double() - ref).abs().max().item() <= 4e-6 * scale


def test_gemm_bounded_follows_device_row_count():
    """escgnn_gemm_tf32x3_bounded: with a device-side row count, output tiles past it are left untouched (rows_dim 1) and
    k-blocks past it are skipped (rows_dim 2); everything before the bound equals the unbounded product."""
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device='cuda').manual_seed(7)
    M, N, K, rows = 1000, 288, 256, 530
    A = torch.randn(M, K, device='cuda', generator=g); B = torch.randn(N, K, device='cuda', generator=g)
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    C = torch.full((M, N), 5.0, device='cuda')
    _lib.check(L.escgnn_gemm_tf32x3_bounded(_p(A), K, 0, _p(B), K, 0, _p(C), N, None, M, N, K, 0, None, 0, _p(d_rows), 1, _st()), 'gemm_tf32x3')
    ref = (A.double() @ B.double().t())
    scale = (A.double().abs() @ B.double().abs().t()).max().item()
    assert (C[:rows].double() - ref[:rows]).abs().max().item() <= 4e-6 * scale
    first_skipped = (rows + 127) // 128 * 128
    assert float(C[first_skipped:].min()) == 5.0 and float(C[first_skipped:].max()) == 5.0      # tiles past the bound: untouched
    # rows_dim 2: wgrad shape, K = row capacity, rows beyond the count hold garbage that must not contribute
    R, cap, n_out, n_in = 700, 1024, 256, 288
    dY = torch.randn(cap, n_out, device='cuda', generator=g); X = torch.randn(cap, n_in, device='cuda', generator=g)
    d_rows2 = torch.tensor([R], dtype=torch.int32, device='cuda')
    dY[((R + 31) // 32) * 32:] = float('nan')                # whole k-blocks past the bound may hold anything
    dY[R:((R + 31) // 32) * 32] = 0.0                        # inside the last k-block the engine keeps tail rows zero
    for mode in (0, 2):
        dW = torch.zeros(n_out, n_in, device='cuda')
        ws = torch.empty(max(L.escgnn_gemm_workspace_floats(n_out, n_in, cap), 1), device='cuda')
        _lib.check(L.escgnn_gemm_tf32x3_bounded(_p(dY), n_out, 1, _p(X), n_in, 1, _p(dW), n_in, None, n_out, n_in, cap, mode, _p(ws),
                                                ws.numel(), _p(d_rows2), 2, _st()), 'gemm_tf32x3')
        ref2 = dY[:R].double().t() @ X[:R].double()
        sc2 = (dY[:R].double().abs().t() @ X[:R].double().abs()).max().item()
        assert torch.isfinite(dW).all() and (dW.double() - ref2).abs().max().item() <= 4e-6 * sc2, mode


@pytest.mark.parametrize('M,N,K', [(256, 10, 2500), (1, 256, 2557), (10, 256, 11063)])
def test_gemm_simple_atomic_split_k(M, N, K):
    """escgnn_gemm_simple with accumulate = 2 (odd-shaped weight gradients of the count variant): K spread over slices that add
    into C atomically; equals C0 + A B^T."""
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device='cuda').manual_seed(M * 7 + N)
    A = torch.randn(M, K, device='cuda', generator=g); B = torch.randn(N, K, device='cuda', generator=g)
    As, Bs = A.t().contiguous(), B.t().contiguous()              # both MN-major, as in wgrad
    C0 = torch.randn(M, N, device='cuda', generator=g)
    C = C0.clone()
    _lib.check(L.escgnn_gemm_simple(_p(As), As.stride(0), 1, _p(Bs), Bs.stride(0), 1, _p(C), N, None, M, N, K, 2, _st()), 'gemm_simple')
    ref = A.double() @ B.double().t() + C0.double()
    scale = (A.double().abs() @ B.double().abs().t()).max().item()
    assert (C.double() - ref).abs().max().item() <= 2e-6 * scale
