"""Back-to-back timing of the BatchNorm+activation kernels at engine shapes (CUDA graph of 40 launches, L2-warm)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esc_gnn_b200 import _lib
L = _lib.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(tag, rows, cap, C, cluster):
    L.escgnn_set_cluster_bn(int(cluster))
    x = torch.randn(cap, C, device='cuda'); y = torch.empty_like(x); dy = torch.randn_like(x); dx = torch.empty_like(x)
    g = torch.ones(C, device='cuda'); b = torch.zeros(C, device='cuda'); rm = torch.zeros(C, device='cuda'); rv = torch.ones(C, device='cuda')
    mean = torch.zeros(C, device='cuda'); rstd = torch.ones(C, device='cuda'); dg = torch.zeros(C, device='cuda'); db = torch.zeros(C, device='cuda')
    part = torch.zeros(L.escgnn_dense_partial_floats(cap, C), device='cuda')
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    f = lambda: _lib.check(L.escgnn_bn_act_fwd(P(x), C, P(g), P(b), P(rm), P(rv), P(mean), P(rstd), P(part), 2, 1e-5, 0.1, 1, P(d_rows), cap, C, P(y), C, st()), 'bn_act_fwd')
    bw = lambda: _lib.check(L.escgnn_bn_act_bwd(P(x), C, P(dy), C, None, 0, P(mean), P(rstd), P(g), P(b), 2, 1, P(part), P(d_rows), cap, C, P(dg), P(db), P(dx), C, st()), 'bn_act_bwd')
    for name, fn in (('fwd', f), ('bwd', bw)):
        for _ in range(3): fn()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(40): fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); e.record(); torch.cuda.synchronize()
        print('%-4s %s rows=%6d C=%4d cluster=%d  %6.2f us per call' % (tag, name, rows, C, cluster, a.elapsed_time(e) * 1e3 / 40))
for cluster in (1, 0):
    run('N', 5906, 6302, 256, cluster)
    run('E', 12092, 12847, 256, cluster)
    run('B', 256, 256, 256, cluster)
def run_act(tag, rows, cap, C):
    x = torch.randn(cap, C, device='cuda'); y = torch.empty_like(x)
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    fn = lambda: _lib.check(L.escgnn_act_fwd(P(x), C, 1, P(d_rows), cap, C, P(y), C, st()), 'act_fwd')
    for _ in range(3): fn()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(40): fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); e.record(); torch.cuda.synchronize()
    print('%-4s act_fwd (plain elementwise kernel, launch floor) rows=%6d  %6.2f us per call' % (tag, rows, a.elapsed_time(e) * 1e3 / 40))
run_act('B', 256, 256, 256)
run_act('N', 5906, 6302, 256)
