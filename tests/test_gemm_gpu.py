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


# ---------------------------------------------------------------------------------------------------------------
# fused Linear -> BatchNorm(training) -> activation (one launch each way) against fp64 torch
def _bn_ref(y, gamma, beta, eps, act):
    mean = y.mean(0)
    var = y.var(0, unbiased=False)
    xh = (y - mean) / torch.sqrt(var + eps)
    z = xh * gamma + beta
    o = torch.relu(z) if act == 1 else torch.nn.functional.elu(z) if act == 2 else z
    return o, mean, var


@pytest.mark.parametrize('act', [1, 2])
@pytest.mark.parametrize('rows_cap,rows,n_out,k_in', [(6302, 5906, 256, 256), (9400, 9300, 256, 256), (700, 700, 300, 600), (130, 1, 64, 32),
                                                      (1000, 513, 96, 40), (2048, 0, 256, 64)])
def test_linear_bn_act_fwd_matches_fp64(act, rows_cap, rows, n_out, k_in):
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    assert L.escgnn_linear_bn_fusable(rows_cap, n_out, k_in) == 1, L.escgnn_linear_bn_resident_ctas(n_out, rows_cap, 0)
    g = torch.Generator(device='cuda').manual_seed(rows_cap + n_out)
    x = torch.randn(rows_cap, k_in, device='cuda', generator=g)
    x[rows:] = 0
    w = torch.randn(n_out, k_in, device='cuda', generator=g) / k_in ** 0.5
    bias = torch.randn(n_out, device='cuda', generator=g)
    gamma = 1 + 0.1 * torch.randn(n_out, device='cuda', generator=g)
    beta = 0.1 * torch.randn(n_out, device='cuda', generator=g)
    rm0, rv0 = torch.randn(n_out, device='cuda', generator=g), 1 + torch.rand(n_out, device='cuda', generator=g)
    rm, rv = rm0.clone(), rv0.clone()
    mean, rstd = torch.zeros(n_out, device='cuda'), torch.zeros(n_out, device='cuda')
    ld = n_out + 32                                       # output is a column slice of a wider buffer
    y = torch.full((rows_cap, n_out), float('nan'), device='cuda')
    out = torch.full((rows_cap, ld), float('nan'), device='cuda')
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    ws_n = L.escgnn_linear_bn_workspace_floats(rows_cap, n_out)
    ws = torch.zeros(ws_n, device='cuda')
    for rep in range(2):                                  # second launch: the barrier tickets were re-armed
        if rep:
            rm.copy_(rm0); rv.copy_(rv0)
        _lib.check(L.escgnn_linear_bn_act_fwd(_p(x), k_in, _p(w), k_in, _p(bias), rows_cap, n_out, k_in, _p(d_rows), _p(gamma), _p(beta),
                                              _p(rm), _p(rv), _p(mean), _p(rstd), act, 1e-5, 0.1, _p(y), n_out, _p(out[:, 16:]), ld,
                                              _p(ws), ws_n, _st()), 'linear_bn_act_fwd')
        torch.cuda.synchronize()
        assert int(ws[:64].view(torch.int32).abs().sum()) == 0
        assert torch.equal(out[rows:, 16:16 + n_out], torch.zeros_like(out[rows:, 16:16 + n_out]))
        assert torch.isnan(out[:, :16]).all() and torch.isnan(out[:, 16 + n_out:]).all()
        if rows == 0:
            continue
        yr = x[:rows].double() @ w.double().t() + bias.double()
        scale = (x[:rows].double().abs() @ w.double().abs().t()).max().item() + 1.0
        assert (y[:rows].double() - yr).abs().max().item() <= 4e-6 * scale
        o, m, v = _bn_ref(yr, gamma.double(), beta.double(), 1e-5, act)
        if rows > 1:
            assert (out[:rows, 16:16 + n_out].double() - o).abs().max().item() <= 2e-4 * max(o.abs().max().item(), 1.0)
            torch.testing.assert_close(mean.double(), m, rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(rstd.double(), 1 / torch.sqrt(v + 1e-5), rtol=2e-4, atol=1e-5)
            ub = v * rows / (rows - 1)
            torch.testing.assert_close(rv.double(), 0.9 * rv0.double() + 0.1 * ub, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(rm.double(), 0.9 * rm0.double() + 0.1 * m, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('act', [1, 2])
@pytest.mark.parametrize('rows_cap,rows,n_in,n_out,bn_cols', [(6302, 5906, 256, 256, 256), (6200, 6092, 288, 1056, 256),
                                                              (700, 650, 600, 300, 600), (300, 77, 64, 96, 32)])
def test_linear_bn_act_bwd_matches_fp64(act, rows_cap, rows, n_in, n_out, bn_cols):
    """dX = BN'(act'(.)) applied to dY W in the dgrad epilogue; columns >= bn_cols are the plain product."""
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    assert L.escgnn_linear_bn_fusable(rows_cap, n_in, n_out) == 1
    g = torch.Generator(device='cuda').manual_seed(rows_cap + n_in)
    dy = torch.randn(rows_cap, n_out, device='cuda', generator=g)
    dy[rows:] = 0
    w = torch.randn(n_out, n_in, device='cuda', generator=g) / n_out ** 0.5
    x = torch.randn(rows_cap, bn_cols, device='cuda', generator=g) * 2 + 0.5
    gamma = 1 + 0.1 * torch.randn(bn_cols, device='cuda', generator=g)
    beta = 0.1 * torch.randn(bn_cols, device='cuda', generator=g)
    x64 = x[:rows].double().requires_grad_(True)
    mean = x64.mean(0)
    var = x64.var(0, unbiased=False)
    o, _, _ = _bn_ref(x64, gamma.double(), beta.double(), 1e-5, act)
    d_act = dy[:rows].double() @ w.double()
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    o2, _, _ = _bn_ref(x64, g64, b64, 1e-5, act)
    (o2 * d_act[:, :bn_cols]).sum().backward()
    mean32, rstd32 = mean.detach().float().contiguous(), (1 / torch.sqrt(var.detach() + 1e-5)).float().contiguous()
    dx = torch.full((rows_cap, n_in), float('nan'), device='cuda')
    dgamma, dbeta = torch.zeros(bn_cols, device='cuda'), torch.zeros(bn_cols, device='cuda')
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    ws_n = L.escgnn_linear_bn_workspace_floats(rows_cap, n_in)
    ws = torch.zeros(ws_n, device='cuda')
    for rep in range(2):
        _lib.check(L.escgnn_linear_bn_act_bwd(_p(dy), n_out, _p(w), n_in, rows_cap, n_in, n_out, _p(d_rows), _p(x), bn_cols, _p(mean32),
                                              _p(rstd32), _p(gamma), _p(beta), act, bn_cols, _p(dgamma), _p(dbeta), _p(dx), n_in, _p(ws),
                                              ws_n, _st()), 'linear_bn_act_bwd')
        torch.cuda.synchronize()
        assert int(ws[:64].view(torch.int32).abs().sum()) == 0
        assert torch.equal(dx[rows:], torch.zeros_like(dx[rows:]))
        sc = d_act.abs().max().item()
        assert (dx[:rows, :bn_cols].double() - x64.grad).abs().max().item() <= 3e-4 * sc
        if bn_cols < n_in:
            assert (dx[:rows, bn_cols:].double() - d_act[:, bn_cols:]).abs().max().item() <= 1e-5 * sc
        torch.testing.assert_close(dgamma.double(), g64.grad, rtol=2e-4, atol=2e-4 * g64.grad.abs().max().item())
        torch.testing.assert_close(dbeta.double(), b64.grad, rtol=2e-4, atol=2e-4 * b64.grad.abs().max().item())


def test_linear_bn_refuses_grids_that_do_not_fit_the_machine():
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    assert L.escgnn_linear_bn_fusable(128 * 400, 256, 256) == 0
    rows_cap, n, k = 128 * 400, 256, 64
    x = torch.zeros(rows_cap, k, device='cuda')
    w = torch.zeros(n, k, device='cuda')
    v = [torch.zeros(n, device='cuda') for _ in range(6)]
    out = torch.zeros(rows_cap, n, device='cuda')
    ws_n = L.escgnn_linear_bn_workspace_floats(rows_cap, n)
    ws = torch.zeros(ws_n, device='cuda')
    rc = L.escgnn_linear_bn_act_fwd(_p(x), k, _p(w), k, None, rows_cap, n, k, None, _p(v[0]), _p(v[1]), _p(v[2]), _p(v[3]), _p(v[4]),
                                    _p(v[5]), 1, 1e-5, 0.1, None, 0, _p(out), n, _p(ws), ws_n, _st())
    assert rc == -2


# ---------------------------------------------------------------------------------------------------------------
# the drain kernel: fp32 accumulation outside the tensor core (escgnn_gemm_set_drain)
@pytest.mark.parametrize('a_mn,b_mn', [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize('M,N,K', [(256, 256, 256), (1000, 288, 256), (300, 160, 96), (6302, 256, 256), (12800, 96, 1056)])
def test_drain_gemm_matches_fp64_and_beats_in_tensor_core_accumulation(a_mn, b_mn, M, N, K):
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    if (a_mn and M % 4) or (b_mn and N % 4) or (not a_mn and K % 4) or (not b_mn and K % 4):
        pytest.skip('pitch not 16-byte aligned for this major')
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    A = torch.randn(M, K, device='cuda', generator=g).abs() + 0.1        # positive operands: the accumulation bias shows as a relative error
    B = torch.randn(N, K, device='cuda', generator=g).abs() + 0.1
    bias = torch.randn(N, device='cuda', generator=g)
    ref = A.double() @ B.double().t()
    was = L.escgnn_gemm_set_drain(2)
    try:
        check(A, B, a_mn, b_mn, bias=bias)
        C0 = torch.randn(M, N, device='cuda', generator=g)
        check(A, B, a_mn, b_mn, C0=C0)
        drained = run_gemm(A, B, a_mn, b_mn)
        L.escgnn_gemm_set_drain(0)
        plain = run_gemm(A, B, a_mn, b_mn)
    finally:
        L.escgnn_gemm_set_drain(was)
    e_d = ((drained.double() - ref) / ref).abs().mean().item()
    e_p = ((plain.double() - ref) / ref).abs().mean().item()
    assert e_d <= 4e-7, e_d                                # fp32-GEMM quality (cuBLAS fp32: ~2e-7 on these shapes)
    if K >= 256:
        assert e_d < 0.5 * e_p, (e_d, e_p)                 # the long in-tensor-core chains are what the drain removes


# ---------------------------------------------------------------------------------------------------------------
# the "wide" plan: one CTA per 128 x 256 tile (escgnn_gemm_set_wide) for 256-column outputs with 64..148 row tiles
@pytest.mark.parametrize('b_mn', [False, True])
@pytest.mark.parametrize('M,N,K', [(12092, 256, 256), (12092, 256, 1056), (9700, 256, 288), (18900, 256, 64), (9000, 512, 96)])
def test_wide_tile_gemm_matches_fp64_and_the_128_wide_tiles(b_mn, M, N, K):
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    A = torch.randn(M, K, device='cuda', generator=g)
    B = torch.randn(N, K, device='cuda', generator=g)
    bias = torch.randn(N, device='cuda', generator=g)
    was = L.escgnn_gemm_set_wide(1)
    try:
        check(A, B, False, b_mn, bias=bias)
        check(A, B, False, b_mn, C0=torch.randn(M, N, device='cuda', generator=g))
        wide = run_gemm(A, B, False, b_mn, bias=bias)
        L.escgnn_gemm_set_wide(0)
        narrow = run_gemm(A, B, False, b_mn, bias=bias)
    finally:
        L.escgnn_gemm_set_wide(was)
    # same products, same k order inside a tile: the two plans agree to rounding of the in-tensor-core accumulation
    scale = (A.abs().double() @ B.abs().double().t()).max().item()
    assert (wide.double() - narrow.double()).abs().max().item() <= 4e-6 * scale
