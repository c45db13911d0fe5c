"""Error of the three GEMM roles of ops.Linear on REAL tensors of the model (cancellation-heavy), against fp64:
   ||C - C64|| / ||C64|| for forward, dgrad, wgrad of a few Linears of the zinc_cfg2 case, product kernel vs cuBLAS fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from esc_gnn_b200 import ops, _lib
from tests import model_util as MU
from tests.test_model_gpu import build_product_model, product_batch

torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else 'zinc_cfg2'
variant, config, count, kw = MU.MODEL_CASES[name]
m = build_product_model(variant, kw).cuda()
sd = MU.det_state(m.state_dict(), seed=1234)
m.load_state_dict({k: v.cuda() for k, v in sd.items()})
m.train()
cap = {}
old = ops.Linear.forward


def fwd(self, x):
    y = F.linear(x, self.weight, self.bias)
    key = [k for k, mod in m.named_modules() if mod is self][0]
    cap[key] = dict(x=x.detach().clone(), w=self.weight.detach().clone(), b=self.bias.detach().clone())
    y.register_hook(lambda g, key=key: cap[key].__setitem__('dy', g.detach().clone()))
    return y


ops.Linear.forward = fwd
b = product_batch(config, 100, count)
MU.loss_fn(variant, m(b), b.y).backward()
ops.Linear.forward = old
rel = lambda a, r: float((a.double() - r).norm() / r.norm())
for mode in (2, 0):
    _lib.lib().escgnn_gemm_set_drain(mode)
    print('---- drain mode', mode)
    for key in ('convs.2.lin', 'convs.1.nn.0', 'convs.1.nn.4', 'z_embedding.3', 'conv1.lin', 'lin1'):
        c = cap[key]
        x, w, bias, dy = c['x'], c['w'], c['b'], c['dy']
        y64 = x.double() @ w.double().t() + bias.double()
        dx64 = dy.double() @ w.double()
        dw64 = dy.double().t() @ x.double()
        xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = bias.clone().requires_grad_(True)
        y = ops._LinearFn.apply(xr, wr, br)
        y.backward(dy)
        yt = F.linear(x, w, bias); dxt = dy @ w; dwt = dy.t() @ x
        print('%-14s [%6d x %4d -> %4d]  fwd %.2e (cublas %.2e)  dgrad %.2e (%.2e)  wgrad %.2e (%.2e)  bias-grad %.2e' % (
            key, x.size(0), x.size(1), w.size(0), rel(y, y64), rel(yt, y64), rel(xr.grad, dx64), rel(dxt, dx64), rel(wr.grad, dw64), rel(dwt, dw64),
            rel(br.grad, dy.double().sum(0))))
