"""ego_rd: general solver with / without pendant-tree peeling, and with the cycle-space fast path in front of it (the default):
ms per 8192 graphs of the config shapes (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esc_gnn_b200 import _lib, synth
from esc_gnn_b200.transform import encode_batch
L = _lib.lib()
for config, G in ((2, 8192), (4, 8192), (1, 8192)):
    fl = synth.ENCODER_FLAGS[config]
    src, dst, eptr, nptr = synth.make_batch_arrays(config, 0, 2048)
    import numpy as np
    reps = G // 2048
    s, d = torch.as_tensor(np.tile(src, reps)).cuda(), torch.as_tensor(np.tile(dst, reps)).cuda()
    ep = torch.as_tensor(np.concatenate([[0], np.cumsum(np.tile(np.diff(eptr), reps))]))
    npt = torch.as_tensor(np.concatenate([[0], np.cumsum(np.tile(np.diff(nptr), reps))]))
    line = 'cfg%d h=%d loops=%d  %d graphs:' % (config, fl['h'], fl['self_loop'], G)
    for peel, fast in ((0, 0), (1, 0), (1, 1)):
        L.escgnn_set_rd_peel(peel)
        L.escgnn_set_rd_fast(fast)
        tm = {}
        for _ in range(3):
            encode_batch(s, d, ep, npt, fl['h'], True, fl['self_loop'], expand=False)
        tm = {}
        for _ in range(5):
            encode_batch(s, d, ep, npt, fl['h'], True, fl['self_loop'], expand=False, timings=tm)
        torch.cuda.synchronize()
        ms = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in tm.items()}
        line += '   peel=%d fast=%d ego_rd %.3f ms ego_encode %.3f ms' % (peel, fast, ms.get('ego_rd', 0), ms.get('ego_encode', 0))
    L.escgnn_set_rd_peel(1)
    L.escgnn_set_rd_fast(1)
    print(line, flush=True)
