"""Are the large per-tensor gradient differences (vs fp64) ReLU kink crossings inside the GINE message relu(x_src + lin(z))?
Counts, per layer, the message entries whose sign differs between the product's fp32 forward and the fp64 oracle, and the
smallest |pre-activation| involved."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import model_util as MU
from tests.test_model_gpu import build_product_model, product_batch

torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else 'zinc_cfg2'
variant, config, count, kw = MU.MODEL_CASES[name]


def capture(model, batch, double):
    cap = {}
    hooks = []
    for k, mod in model.named_modules():
        if k == 'conv1' or (k.startswith('convs.') and k.count('.') == 1):
            hooks.append(mod.register_forward_pre_hook(lambda m_, args, k=k: cap.__setitem__(k, (args[0].detach(), args[2].detach(), m_.lin))))
    model(batch)
    for h in hooks:
        h.remove()
    out = {}
    for k, (x, z, lin) in cap.items():
        src = batch.edge_index[0]
        pre = x[src] + torch.nn.functional.linear(z.double(), lin.weight.double(), lin.bias.double()).to(x.dtype) if not double else \
            x[src] + torch.nn.functional.linear(z, lin.weight, lin.bias)
        out[k] = pre.double()
    return out


m = build_product_model(variant, kw).cuda()
sd = MU.det_state(m.state_dict(), seed=1234)
m.load_state_dict({k: v.cuda() for k, v in sd.items()})
m.train()
b = product_batch(config, 100, count)
p32 = capture(m, b, False)
o = MU.build_oracle_model(variant, kw).double()
o.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()})
o = o.cuda().train()
rb = MU.to_double(MU.ref_batch(config, 100, count))
for k, v in list(rb.__dict__.items()):
    if torch.is_tensor(v):
        setattr(rb, k, v.cuda())
p64 = capture(o, rb, True)
for k in sorted(p32):
    a, r = p32[k], p64[k]
    flips = ((a > 0) != (r > 0))
    near = (r.abs() < 1e-5).sum().item()
    print('%-8s entries %9d  sign flips %4d  |pre64| < 1e-5: %5d  max |pre32-pre64| %.2e  min |pre64| at flips %s' % (
        k, a.numel(), int(flips.sum()), near, float((a - r).abs().max()), ('%.2e' % float(r[flips].abs().max())) if flips.any() else '-'))
