"""Where does the product's gradient error (vs the fp64 truth) come from?  Module path with individual kernels swapped for
plain torch fp32 ops, per-tensor ratio  ||g - g64|| / ||g32_reference - g64||  (fixture yardstick).

    python tools/debug_grad_error.py [zinc_cfg2]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from esc_gnn_b200 import ops
from tests import model_util as MU
from tests.test_model_gpu import FIX_M, _fp64_truth, build_product_model, product_batch

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else 'zinc_cfg2'
variant, config, count, kw = MU.MODEL_CASES[name]
g64 = _fp64_truth(name)
keys = [str(k) for k in FIX_M[name + '/grad_keys']]
ref_err = dict(zip(keys, FIX_M[name + '/grad_err32']))


def run(label, linear=None, bn=None):
    old_l, old_b = ops.Linear.forward, ops.BatchNorm1d.forward
    if linear is not None:
        ops.Linear.forward = linear
    if bn is not None:
        ops.BatchNorm1d.forward = bn
    try:
        m = build_product_model(variant, kw).cuda()
        sd = MU.det_state(m.state_dict(), seed=1234)
        m.load_state_dict({k: v.cuda() for k, v in sd.items()})
        m.train()
        b = product_batch(config, 100, count)
        MU.loss_fn(variant, m(b), b.y).backward()
        g = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    finally:
        ops.Linear.forward, ops.BatchNorm1d.forward = old_l, old_b
    rows = []
    for k in keys:
        if variant == 'count' and k.startswith('x_embedding.'):
            continue
        err = float((g[k].double() - g64[k]).norm())
        rows.append((err / max(3 * ref_err[k] + 2e-6 * float(g64[k].norm()), 1e-30), k, err, ref_err[k], float(g64[k].norm())))
    rows.sort(reverse=True)
    print('== %-34s worst err/bound %.2f   median %.2f' % (label, rows[0][0], rows[len(rows) // 2][0]))
    for r in rows[:8]:
        print('     %-30s ratio %6.2f err %.3e ref_err %.3e norm %.3e' % (r[1], r[0], r[2], r[3], r[4]))
    return g


torch_linear = lambda self, x: F.linear(x, self.weight, self.bias)
torch_bn = lambda self, x: F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, self.training, self.momentum, self.eps)
run('product kernels')
from esc_gnn_b200 import _lib
_lib.lib().escgnn_set_cluster_bn(0)
run('product kernels, cluster BN off (stats + apply kernels)')
_lib.lib().escgnn_set_cluster_bn(1)
run('torch Linear (cuBLAS fp32)', linear=torch_linear)
run('torch BatchNorm', bn=torch_bn)
run('torch Linear + torch BatchNorm', linear=torch_linear, bn=torch_bn)
# pure torch fp32 oracle on the GPU (what plain fp32 arithmetic on this device gives)
o = MU.build_oracle_model(variant, kw)
sd = MU.det_state(o.state_dict(), seed=1234)
o.load_state_dict(sd)
o = o.cuda().train()
rb = MU.ref_batch(config, 100, count)
for k, v in list(rb.__dict__.items()):
    if torch.is_tensor(v):
        setattr(rb, k, v.cuda())
MU.loss_fn(variant, o(rb), rb.y).backward()
rows = []
for k, p in o.named_parameters():
    if p.grad is None or (variant == 'count' and k.startswith('x_embedding.')):
        continue
    err = float((p.grad.double() - g64[k]).norm())
    rows.append((err / max(3 * ref_err[k] + 2e-6 * float(g64[k].norm()), 1e-30), k, err, ref_err[k]))
rows.sort(reverse=True)
print('== pure torch fp32 oracle on the GPU   worst err/bound %.2f median %.2f' % (rows[0][0], rows[len(rows) // 2][0]))
for r in rows[:5]:
    print('     %-30s ratio %6.2f err %.3e ref_err %.3e' % (r[1], r[0], r[2], r[3]))
