"""Short encoder run for ncu (config 2, 8192 ZINC-shaped graphs, 3 iterations)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esc_gnn_b200 import synth  # noqa: E402
from esc_gnn_b200.transform import encode_batch  # noqa: E402
from tools.enc_bench import tiled  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fl = synth.ENCODER_FLAGS[cfg]
src, dst, eptr, nptr = tiled(cfg, 256, 8192)
ds, dd = torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda()
for _ in range(3):
    r = encode_batch(ds, dd, torch.as_tensor(eptr), torch.as_tensor(nptr), **fl)
torch.cuda.synchronize()
print('ok', r.nnz)
