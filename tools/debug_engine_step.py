import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import model_util as MU
from tests.test_model_gpu import _engine_for
torch.backends.cuda.matmul.allow_tf32 = False
variant, config, count, kw = MU.MODEL_CASES['zinc']
eng, model, raw = _engine_for(variant, config, count, kw, use_graph=False)
eng.load(raw)
with torch.no_grad():
    eng._encode_and_index()
    eng.opt.grad.zero_()
    for f in eng.fwd:
        f()
    torch.cuda.synchronize()
    B = eng.debug_buffers
    snap = {k: v.clone() for k, v in B.items() if not k.startswith('d')}
    for i, b in enumerate(reversed(eng.bwd)):
        b()
        torch.cuda.synchronize()
        for k, v in snap.items():
            if not torch.equal(B[k], v):
                print('bwd op', i, 'changed', k, (B[k] - v).abs().max().item())
                snap[k] = B[k].clone()
print('done', len(eng.bwd))
