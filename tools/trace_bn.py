"""Where a node-level BatchNorm of the training chain spends its time: %globaltimer stamps from inside the register-resident
cluster kernels (escgnn_bn_set_trace) for a chain  bn_fwd -> bn_bwd -> bn_fwd ...  replayed as a CUDA graph with programmatic
dependent launches.  Microseconds after the CTA's dependency wait returned, averaged over CTAs and launches."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esc_gnn_b200 import _lib
L = _lib.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
NAMES = ['CTA start', 'dependency wait returned', 'tile loaded + summed', 'cluster reduction done', 'outputs stored', 'cluster released']


def chain(rows, C, reps=12):
    x = torch.randn(rows, C, device='cuda'); y = torch.empty_like(x); dy = torch.randn_like(x); dx = torch.empty_like(x)
    g = torch.ones(C, device='cuda'); b = torch.zeros(C, device='cuda'); rm = torch.zeros(C, device='cuda'); rv = torch.ones(C, device='cuda')
    mean = torch.zeros(C, device='cuda'); rstd = torch.ones(C, device='cuda'); dg = torch.zeros(C, device='cuda'); db = torch.zeros(C, device='cuda')
    part = torch.zeros(int(L.escgnn_dense_partial_floats(rows, C)), device='cuda')
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    ctas = 8 * ((C + 7) // 8)
    tr = torch.zeros(2 * reps, ctas * 6, dtype=torch.int64, device='cuda')

    def run():
        for i in range(reps):
            L.escgnn_bn_set_trace(ctypes.c_void_p(tr[2 * i].data_ptr()))
            _lib.check(L.escgnn_bn_act_fwd(P(x), C, P(g), P(b), P(rm), P(rv), P(mean), P(rstd), P(part), 2, 1e-5, 0.1, 1, P(d_rows), rows, C,
                                           P(y), C, st()), 'fwd')
            L.escgnn_bn_set_trace(ctypes.c_void_p(tr[2 * i + 1].data_ptr()))
            _lib.check(L.escgnn_bn_act_bwd(P(x), C, P(dy), C, None, 0, P(mean), P(rstd), P(g), P(b), 2, 1, P(part), P(d_rows), rows, C,
                                           P(dg), P(db), P(dx), C, st()), 'bwd')
        L.escgnn_bn_set_trace(None)
    # NOTE: the trace pointer is a device symbol read at RUN time: a captured graph would see only the last value, so the chain
    # runs eagerly on one stream (launches are still programmatic-dependent)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    t = tr.view(reps, 2, ctas, 6)[2:].double() / 1e3
    for which, nm in ((0, 'forward'), (1, 'backward')):
        tt = t[:, which]
        rel = tt - tt[:, :, 1:2]
        print('%s  rows=%d C=%d (%d CTAs)' % (nm, rows, C, ctas))
        for i, n in enumerate(NAMES):
            print('   %-26s mean %7.2f   max over CTAs %7.2f' % (n, rel[:, :, i].mean().item(), rel[:, :, i].max(dim=1).values.mean().item()))
        print('   span (first wait return -> last release) %.2f us' % (tt[:, :, 5].max(dim=1).values - tt[:, :, 1].min(dim=1).values).mean().item())


chain(5906, 256)
chain(12092, 256)
chain(640, 256)
