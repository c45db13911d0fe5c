"""GPU: row-wise dense kernels of the static engine (BN+act fwd/bwd, colsum, embedding, loss, Adam) vs torch fp32."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize('rows,cap,C,act', [(300, 384, 256, 'elu'), (48, 48, 256, 'relu'), (1000, 1500, 10, 'relu'),
                                            (257, 400, 300, 'none'), (5, 130, 32, 'elu'), (70000, 70100, 64, 'relu')])
def test_bn_act_forward_backward_match_torch(rows, cap, C, act):
    from esc_gnn_b200 import _lib
    from esc_gnn_b200.engine import ACT
    L = _lib.lib()
    g = torch.Generator(device='cuda').manual_seed(rows + C)
    ld = C + 8                                            # strided views (column slices of wider buffers)
    xbuf = torch.randn(cap, ld, device='cuda', generator=g) * 2 + 0.5
    x = xbuf[:, :C]
    bn = torch.nn.BatchNorm1d(C).cuda()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.1 * torch.randn(C, device='cuda', generator=g)); bn.bias.copy_(0.1 * torch.randn(C, device='cuda', generator=g))
    ref_bn = torch.nn.BatchNorm1d(C).cuda(); ref_bn.load_state_dict(bn.state_dict())
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    partial = torch.zeros(L.escgnn_dense_partial_floats(cap, C), device='cuda')
    mean, rstd = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    ybuf = torch.full((cap + 200, ld), 7.0, device='cuda'); y = ybuf[:cap, 4:4 + C]     # 200 guard rows past the capacity
    _lib.check(L.escgnn_bn_act_fwd(_p(x), ld, _p(bn.weight), _p(bn.bias), _p(bn.running_mean), _p(bn.running_var), _p(mean),
                                   _p(rstd), _p(partial), ACT[act], bn.eps, bn.momentum, 1, _p(d_rows), cap, C, _p(y), ld,
                                   _st()), 'bn_act_fwd')
    xr = x[:rows].clone().requires_grad_(True)
    f = {'elu': torch.nn.functional.elu, 'relu': torch.relu, 'none': lambda t: t}[act]
    yr = f(ref_bn(xr))
    torch.testing.assert_close(y[:rows], yr, rtol=1e-5, atol=1e-5)
    assert float(y[rows:].abs().max()) == 0.0 if rows < cap else True
    assert float(ybuf[cap:].min()) == 7.0 and float(ybuf[cap:].max()) == 7.0, 'wrote past rows_cap'
    torch.testing.assert_close(bn.running_mean, ref_bn.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn.running_var, ref_bn.running_var, rtol=1e-5, atol=1e-6)
    dy = torch.randn(cap, C, device='cuda', generator=g); dy2 = torch.randn(cap, C, device='cuda', generator=g)
    dgamma, dbeta = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    dxbuf = torch.full((cap + 200, C), 3.0, device='cuda'); dx = dxbuf[:cap]
    _lib.check(L.escgnn_bn_act_bwd(_p(x), ld, _p(dy), C, _p(dy2), C, _p(mean), _p(rstd), _p(bn.weight), _p(bn.bias), ACT[act], 1,
                                   _p(partial), _p(d_rows), cap, C, _p(dgamma), _p(dbeta), _p(dx), C, _st()), 'bn_act_bwd')
    yr.backward(dy[:rows] + dy2[:rows])
    torch.testing.assert_close(dx[:rows], xr.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dgamma, ref_bn.weight.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dbeta, ref_bn.bias.grad, rtol=1e-4, atol=1e-4)
    if rows < cap:
        assert float(dx[rows:].abs().max()) == 0.0
    assert float(dxbuf[cap:].min()) == 3.0 and float(dxbuf[cap:].max()) == 3.0, 'wrote past rows_cap'


@pytest.mark.parametrize('rows,cap,C', [(48, 48, 1), (48, 48, 256), (300, 1000, 288), (1, 200, 7), (70000, 70100, 64)])
def test_colsum_matches_torch(rows, cap, C):
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    x = torch.randn(cap, C, device='cuda')
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    partial = torch.zeros(L.escgnn_dense_partial_floats(cap, C), device='cuda')
    out = torch.zeros(C, device='cuda')
    _lib.check(L.escgnn_colsum(_p(x), C, _p(d_rows), cap, C, _p(partial), _p(out), _st()), 'colsum')
    torch.testing.assert_close(out, x[:rows].sum(0), rtol=1e-5, atol=1e-5 * max(1.0, rows ** 0.5))


def test_embedding_loss_and_adam_match_torch():
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device='cuda').manual_seed(5)
    table = torch.nn.Embedding(100, 32).cuda()
    rows, cap = 70, 96
    idx = torch.randint(0, 100, (cap, ), device='cuda', generator=g)
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    ybuf = torch.full((cap, 40), 9.0, device='cuda'); y = ybuf[:, 8:]
    _lib.check(L.escgnn_embedding_fwd(_p(table.weight), _p(idx), 1, None, _p(d_rows), cap, 32, _p(y), 40, _st()), 'embedding_fwd')
    torch.testing.assert_close(y[:rows], table(idx[:rows]))
    assert float(y[rows:].abs().max()) == 0.0 and float(ybuf[:, :8].min()) == 9.0
    dy = torch.randn(cap, 32, device='cuda', generator=g)
    dt = torch.zeros_like(table.weight)
    _lib.check(L.escgnn_embedding_bwd(_p(dy), 32, _p(idx), 1, None, _p(d_rows), cap, 32, _p(dt), _st()), 'embedding_bwd')
    table(idx[:rows]).backward(dy[:rows])
    torch.testing.assert_close(dt, table.weight.grad, rtol=1e-5, atol=1e-5)
    # losses
    for kind in (0, 1):
        pred = torch.randn(cap, 1, device='cuda', generator=g, requires_grad=True)
        tgt = torch.randn(cap, device='cuda', generator=g) if kind == 0 else (torch.rand(cap, device='cuda', generator=g) < 0.3).float()
        if kind == 1:
            tgt[3] = float('nan')
        loss, dpred = torch.zeros(1, device='cuda'), torch.full((cap, 1), 5.0, device='cuda')
        _lib.check(L.escgnn_loss_fwd_bwd(_p(pred), 1, _p(tgt), kind, _p(d_rows), cap, 1, _p(loss), _p(dpred), 1, _st()), 'loss_fwd_bwd')
        if kind == 0:
            ref = torch.nn.L1Loss()(pred[:rows], tgt[:rows].view(-1, 1))
        else:
            lab = tgt[:rows] == tgt[:rows]
            ref = torch.nn.BCEWithLogitsLoss()(pred[:rows][lab.view(-1)], tgt[:rows][lab].view(-1, 1))
        ref.backward()
        torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(dpred[:rows], pred.grad[:rows], rtol=1e-5, atol=1e-7)
        assert float(dpred[rows:].abs().max()) == 0.0
    # Adam (device-side step counter) vs torch.optim.Adam over 4 steps
    from esc_gnn_b200.optim import FlatAdam
    lin_a, lin_b = torch.nn.Linear(37, 19).cuda(), torch.nn.Linear(37, 19).cuda()
    lin_b.load_state_dict(lin_a.state_dict())
    fa, ta = FlatAdam(lin_a.parameters(), lr=1e-2), torch.optim.Adam(lin_b.parameters(), lr=1e-2)
    xin = torch.randn(64, 37, device='cuda', generator=g)
    for step in range(4):
        fa.zero_grad(); ta.zero_grad()
        lin_a(xin).pow(2).mean().backward(); lin_b(xin).pow(2).mean().backward()
        if step % 2 == 0:
            fa.step_device()
        else:
            fa.t = step; fa.step(); fa.state += 1        # host-scalar variant, same update
        ta.step()
        torch.testing.assert_close(lin_a.weight, lin_b.weight, rtol=1e-5, atol=1e-6)
