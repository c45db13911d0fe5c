"""Where a node-level GEMM of the training chain spends its time: %globaltimer stamps from inside gemm_tf32x3_ts_kernel
(escgnn_gemm_set_trace) for a chain  act -> GEMM -> act -> GEMM ...  replayed as a CUDA graph with programmatic dependent launches.
Times are microseconds after the moment the CTA's dependency wait returned (stamp 2), averaged over CTAs and launches."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esc_gnn_b200 import _lib
L = _lib.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
SLOTS = L.escgnn_gemm_trace_slots()      # 72 with a library built by ESCGNN_NVCC_FLAGS=-DESCGNN_TRACE_KB python -m esc_gnn_b200.build --force
NAMES = ['CTA start', 'prologue done', 'dependency wait returned', 'first stage landed', 'first k-block split', 'last MMA issued',
         'accumulator complete', 'epilogue stored']


def chain(M, N, K, reps=12, b_mn=0, detail=False):
    X = torch.randn(M, K, device='cuda'); W = torch.randn(K, N, device='cuda') if b_mn else torch.randn(N, K, device='cuda')
    Y = torch.empty(M, N, device='cuda'); Z = torch.empty(M, K, device='cuda')
    rows = torch.tensor([M], dtype=torch.int32, device='cuda')
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    tr = torch.zeros(reps, tiles * SLOTS, dtype=torch.int64, device='cuda')
    ws = torch.empty(1 << 20, device='cuda')

    def run():
        for i in range(reps):
            L.escgnn_gemm_set_trace(ctypes.c_void_p(tr[i].data_ptr()))
            _lib.check(L.escgnn_gemm_tf32x3_bounded(P(X), X.stride(0), 0, P(W), W.stride(0), b_mn, P(Y), N, None, M, N, K, 0, P(ws), ws.numel(),
                                                    P(rows), 1 | 4, st()), 'gemm')
            _lib.check(L.escgnn_act_fwd(P(Y), Y.stride(0), 1, P(rows), M, min(N, K), P(Z), Z.stride(0), st()), 'act')
        L.escgnn_gemm_set_trace(None)
    run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    t = tr[2:].view(reps - 2, tiles, SLOTS).double() / 1e3        # us; the first launches warm up
    rel = t - t[:, :, 2:3]
    print('M=%d N=%d K=%d (%d CTAs), B %s-major' % (M, N, K, tiles, 'MN' if b_mn else 'K'))
    for i, nm in enumerate(NAMES):
        print('   %-26s mean %7.2f   max over CTAs %7.2f' % (nm, rel[:, :, i].mean().item(), rel[:, :, i].max(dim=1).values.mean().item()))
    kb = (K + 31) // 32
    if kb <= 16 and detail and SLOTS >= 72:
        print('   per k-block (mean over CTAs): stage free / landed / split / MMAs issued')
        for i in range(kb):
            print('     k-block %2d  %6.2f %6.2f %6.2f %6.2f' % ((i, ) + tuple(rel[:, :, 8 + 4 * i + j].mean().item() for j in range(4))))
    span = (t[:, :, 7].max(dim=1).values - t[:, :, 2].min(dim=1).values).mean().item()
    period = (t[1:, :, 2].min(dim=1).values - t[:-1, :, 2].min(dim=1).values).mean().item()
    gap = (t[1:, :, 2].min(dim=1).values - t[:-1, :, 7].max(dim=1).values).mean().item()
    print('   kernel span (first wait return -> last store) %.2f us; GEMM+act period %.2f us; last store -> next GEMM released %.2f us (the act kernel + 2 boundaries)' % (span, period, gap))


for sw, staged, kbg in ((4, 0, 1), (4, 1, 1), (8, 1, 1), (8, 1, 2)):
    L.escgnn_gemm_set_split_warps(sw)
    L.escgnn_gemm_set_staged_store(staged)
    L.escgnn_gemm_set_kb_groups(kbg)
    print('---- %d splitter / epilogue warps, %s epilogue, %d k-block group(s)' % (sw, 'staged (row-contiguous)' if staged else 'register-row', kbg))
    chain(5906, 256, 256, detail=(kbg == 2))
    chain(5906, 256, 256, b_mn=1)
    chain(640, 256, 256)
L.escgnn_gemm_set_split_warps(8)
L.escgnn_gemm_set_staged_store(1)
L.escgnn_gemm_set_kb_groups(2)
