"""Loss trajectories of two sequential engines and one pipelined engine on the same batches (run-to-run noise vs a real difference)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import model_util as MU
from tests.test_model_gpu import _engine_for
from esc_gnn_b200.pipeline import RawBatch
for name in ('count_h64', 'zinc'):
    variant, config, count, kw = MU.MODEL_CASES[name]
    a, _, raw0 = _engine_for(variant, config, count, kw, True)
    b, _, _ = _engine_for(variant, config, count, kw, True)
    p, _, _ = _engine_for(variant, config, count, kw, True, pipeline=True)
    raws = [raw0] + [RawBatch.synth(config, 100 + 7 * i, count) for i in (1, 2)]
    raws = [r for r in raws if r.num_nodes <= a.c.caps['N'] and r.src.numel() <= a.c.caps['E_in']]
    order = [raws[i % len(raws)] for i in range(9)]
    la = [float(a.step(r).item()) for r in order]
    lb = [float(b.step(r).item()) for r in order]
    assert p.step(order[0]) is None
    lp = [float(p.step(r).item()) for r in order[1:]] + [float(p.drain().item())]
    print(name, len(raws))
    for x, y, z in zip(la, lb, lp):
        print('  %.7f %.7f %.7f   seq-seq %.2e  seq-pipe %.2e' % (x, y, z, abs(x - y) / abs(x), abs(x - z) / abs(x)))
