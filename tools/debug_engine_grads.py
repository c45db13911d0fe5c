import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import model_util as MU
from tests.test_model_gpu import _engine_for, build_product_model, product_batch
torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else 'zinc'
variant, config, count, kw = MU.MODEL_CASES[name]
eng, model, raw = _engine_for(variant, config, count, kw, use_graph=False)
eng.opt.param_groups[0]['lr'] = 0.0
loss_e = float(eng.step(raw).item())
ge = {k: p.grad.clone() for k, p in model.named_parameters()}
ref = build_product_model(variant, kw).cuda()
sd = MU.det_state(ref.state_dict(), seed=1234)
ref.load_state_dict({k: v.cuda() for k, v in sd.items()}); ref.train()
b = product_batch(config, 100, count)
cap = {}
ref.lin1.register_forward_hook(lambda m, i, o: cap.__setitem__('in', i[0].detach().clone()))
ref.lin1.register_full_backward_hook(lambda m, gi, go_: cap.__setitem__('gout', go_[0].detach().clone()))
lm = MU.loss_fn(variant, ref(b), b.y); lm.backward()
gm = {k: p.grad.clone() for k, p in ref.named_parameters()}
# CPU double oracle
o = MU.build_oracle_model(variant, kw).double()
o.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in sd.items()}); o.train()
rb = MU.ref_batch(config, 100, count)
for a in ('x', 'y'):
    t = getattr(rb, a)
    if t.is_floating_point(): setattr(rb, a, t.double())
rb.pos_enc = rb.pos_enc
import oracle.model_ref as R
_old = R.bag_embed
R.bag_embed = lambda w, pi, pe, pb: _old(w, pi, pe.double(), pb)
lo = MU.loss_fn(variant, o(rb).float(), rb.y.float()) if False else torch.nn.L1Loss()(o(rb), rb.y.view(-1, 1))
lo.backward()
go = {k: p.grad.float().cuda() for k, p in o.named_parameters()}
print('loss engine %.7f module %.7f fp64 %.7f' % (loss_e, lm.item(), lo.item()))
print('%-40s %10s %10s %10s' % ('param', 'eng-vs-64', 'mod-vs-64', 'eng-vs-mod'))
for k in gm:
    s = go[k].abs().max().item() + 1e-12
    print('%-40s %10.2e %10.2e %10.2e' % (k, (ge[k] - go[k]).abs().max().item() / s, (gm[k] - go[k]).abs().max().item() / s,
                                        (ge[k] - gm[k]).abs().max().item() / s))
for k in ('lin2.bias', 'lin1.weight', 'lin2.weight'):
    print(k, 'engine', ge[k].flatten()[:6].tolist(), 'module', gm[k].flatten()[:6].tolist())
d = (ge['lin1.weight'] - gm['lin1.weight']).abs()
print('lin1.weight err by column block', [d[:, i * 256:(i + 1) * 256].max().item() for i in range(5)])
print('rows of max err', d.max(1).values.topk(5))

B = eng.debug_buffers
print('lin2.bias engine', ge['lin2.bias'].tolist(), 'module', gm['lin2.bias'].tolist(), 'manual', B['dpred'].sum(0).tolist())
man = B['dp1'].t() @ B['head_in']
print('lin1.weight manual-vs-engine', (man - ge['lin1.weight']).abs().max().item(), 'manual-vs-module', (man - gm['lin1.weight']).abs().max().item())
print('dims', eng.c.dims.tolist(), 'graph_ptr tail', eng.graph_ptr[-3:].tolist())

print('pooled engine-vs-module', (B['head_in'] - cap['in']).abs().max().item(), 'scale', cap['in'].abs().max().item())
print('dp1 engine-vs-module', (B['dp1'] - cap['gout']).abs().max().item(), 'scale', cap['gout'].abs().max().item())
print('dp1 colsum engine', B['dp1'].sum(0)[:4].tolist(), 'module', cap['gout'].sum(0)[:4].tolist())
