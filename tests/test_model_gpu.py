"""GPU: the sm_100a model kernels and the three NestedGIN_eff variants against the torch oracle and the fixtures
produced by the reference's own classes.  Tolerance: 1e-4 relative in fp32 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import model_ref
from tests import model_util as MU
from tests.test_batch_cpu import to_data
from tests.test_model_oracle_cpu import run_case

pytestmark = pytest.mark.gpu
RTOL = 1e-4
from tests.test_model_oracle_cpu import FIX as FIX_M  # noqa: E402


def product_batch(config, start, count):
    from esc_gnn_b200.batch import Batch
    return Batch.from_data_list([to_data(g) for g in MU.graph_dicts(config, start, count)]).to('cuda')


def qm9_product_batch(start, count):
    """The reference's QM9 flow end to end on the product side (run_qm9.py:200-231): raw molecule -> create_subgraphs
    (h=3, rd, self-loops; a Data with `name` keeps pos / node_type, utils_edge_efficient.py:150-151) -> Distance ->
    collation."""
    from esc_gnn_b200 import synth
    from esc_gnn_b200.batch import Batch
    from esc_gnn_b200.data import Data
    from esc_gnn_b200.distance import Distance
    from esc_gnn_b200.transform import create_subgraphs
    out = []
    for i in range(start, start + count):
        g = synth.make_graph(6, i)
        d = Data(x=torch.as_tensor(g['x']), edge_index=torch.as_tensor(g['edge_index']), edge_attr=torch.as_tensor(g['edge_attr']),
                 y=torch.as_tensor(g['y']).view(1), pos=torch.as_tensor(g['pos']), name='gdb_%d' % i,
                 node_type=torch.as_tensor(g['node_type']))
        d = create_subgraphs(d, 3, node_label='hop', use_rd=True, subgraph_pretransform=None, self_loop=True)
        out.append(Distance(norm=True, relative_pos=False, squared=False)(d.to('cuda')))
    return Batch.from_data_list(out).to('cuda')


def build_product_model(variant, kw):
    from esc_gnn_b200 import graphcount_model, ogb_model, qm9_model, synth, zinc_model
    if variant == 'qm9':
        class _QM9(object):
            num_features = synth.QM9_FEATURES
        return qm9_model.NestedGIN_eff(_QM9(), kw['num_layers'])
    if variant == 'kgin':
        from esc_gnn_b200 import kernel_gin_model

        class _TU(object):
            num_features, num_classes = synth.KGIN_FEATURES, synth.KGIN_CLASSES
        return kernel_gin_model.NestedGIN_eff(_TU(), kw['num_layers'], kw['hidden'], use_rd=True, dropout=0)
    if variant == 'count':
        return graphcount_model.NestedGIN_eff(None, kw['num_layers'], kw['hidden'], use_rd=True, graph_pred=False,
                                              dropout=0, edge_nest=True, use_cycle=True)
    if variant == 'zinc':
        return zinc_model.NestedGIN_eff(None, kw['num_layers'])
    return ogb_model.GNN('ogbg-molhiv', kw['num_tasks'], num_layer=kw['num_layer'], emb_dim=kw['emb_dim'],
                         gnn_type='gin_eff', virtual_node=kw['virtual_node'], residual=kw['residual'],
                         drop_ratio=kw['drop_ratio'])


@pytest.mark.parametrize('name', list(MU.MODEL_CASES))
def test_models_match_reference_class_fixtures(name):
    variant, config, count, kw = MU.MODEL_CASES[name]
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    model = build_product_model(variant, kw).cuda()
    batch = qm9_product_batch(100, count) if variant == 'qm9' else product_batch(config, 100, count)
    run_case(model, variant, batch, name, rtol=RTOL, atol=1e-5)


def test_ops_against_torch_oracle():
    from esc_gnn_b200 import ops
    b = product_batch(2, 700, 40)
    idx = ops.graph_index(b)
    E, N = b.edge_index.size(1), b.x.size(0)
    g = torch.Generator(device='cuda').manual_seed(3)
    for H in (256, 300, 64):
        W = torch.randn(1800, H, device='cuda', generator=g, requires_grad=True)
        W2 = W.detach().clone().requires_grad_(True)
        out = ops.bag_embed(W, b.pos_index, b.pos_enc, idx)
        ref = model_ref.bag_embed(W2, b.pos_index, b.pos_enc, b.pos_batch)
        torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-4)
        go = torch.randn(E, H, device='cuda', generator=g)
        out.backward(go); ref.backward(go)
        torch.testing.assert_close(W.grad, W2.grad, rtol=1e-4, atol=1e-3)
    for C in (256, 32, 10, 300):
        x = torch.randn(N, C, device='cuda', generator=g, requires_grad=True)
        e = torch.randn(E, C, device='cuda', generator=g, requires_grad=True)
        eps = torch.tensor([0.3], device='cuda', requires_grad=True)
        x2, e2, eps2 = (t.detach().clone().requires_grad_(True) for t in (x, e, eps))
        out = ops.gine_aggregate(x, e, eps, idx)
        msg = (x2[b.edge_index[0]] + e2).relu()
        ref = torch.zeros_like(x2).index_add(0, b.edge_index[1], msg) + (1 + eps2) * x2
        torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
        go = torch.randn(N, C, device='cuda', generator=g)
        out.backward(go); ref.backward(go)
        torch.testing.assert_close(x.grad, x2.grad, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(e.grad, e2.grad, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(eps.grad, eps2.grad, rtol=1e-4, atol=1e-3)
    x = torch.randn(N, 96, device='cuda', generator=g, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    for mean in (False, True):
        out = (ops.global_mean_pool if mean else ops.global_add_pool)(x, idx)
        ref = (model_ref.global_mean_pool if mean else model_ref.global_add_pool)(x2, b.batch)
        torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
        go = torch.randn_like(out)
        x.grad = None; x2.grad = None
        out.backward(go); ref.backward(go)
        torch.testing.assert_close(x.grad, x2.grad, rtol=1e-5, atol=1e-6)


def test_csr_build_is_sorted_and_complete():
    from esc_gnn_b200 import ops
    b = product_batch(1, 50, 30)
    idx = ops.graph_index(b)
    dst = b.edge_index[1].cpu().numpy(); src = b.edge_index[0].cpu().numpy()
    for keys, ptr, perm in ((dst, idx.dst_ptr, idx.dst_perm), (src, idx.src_ptr, idx.src_perm)):
        ptr, perm = ptr.cpu().numpy(), perm.cpu().numpy()
        assert sorted(perm.tolist()) == list(range(len(keys)))
        for i in range(len(ptr) - 1):
            seg = perm[ptr[i]:ptr[i + 1]]
            assert (keys[seg] == i).all() and (np.diff(seg) > 0).all()
    assert int(idx.err[0]) == 0
    assert np.array_equal(idx.rec_ptr.cpu().numpy(),
                          np.searchsorted(b.pos_batch.cpu().numpy(), np.arange(len(dst) + 1)))


def test_distance_transform_matches_reference_formula():
    from esc_gnn_b200.data import Data
    from esc_gnn_b200.distance import Distance
    g = torch.Generator().manual_seed(0)
    pos = torch.randn(40, 3, generator=g)
    ei = torch.randint(0, 40, (2, 200), generator=g)
    ea = torch.randn(200, 4, generator=g)
    for dev in ('cpu', 'cuda'):
        for kw in (dict(), dict(norm=False), dict(squared=True), dict(relative_pos=True), dict(max_value=3.0),
                   dict(cat=False)):
            d = Data(x=torch.ones(40, 1), edge_index=ei.to(dev), edge_attr=ea.to(dev), pos=pos.to(dev))
            out = Distance(**kw)(d).edge_attr.cpu()
            row, col = ei
            dist = ((pos[col] - pos[row]) ** 2).sum(1).view(-1, 1) if kw.get('squared') else \
                torch.norm(pos[col] - pos[row], p=2, dim=-1).view(-1, 1)
            if kw.get('norm', True):
                dist = dist / (dist.max() if kw.get('max_value') is None else kw['max_value'])
            want = torch.cat([ea, dist], -1) if kw.get('cat', True) else dist
            if kw.get('relative_pos'):
                want = torch.cat([want, pos[col] - pos[row]], -1)
            torch.testing.assert_close(out, want, rtol=1e-6, atol=1e-6)


def test_distance_transform_matches_reference_fixtures():
    """D1 pinned: outputs of the unmodified /root/reference/distance.py (tests/golden/make_golden_batch.py -> batch.npz), every
    constructor option, 1-D / missing edge_attr, and the original_* branch (distance.py:49-63); CPU-resident and CUDA-resident
    data.  fp32 tolerance 1e-6 relative (sqrt / division rounding; the reference is plain torch fp32)."""
    import os
    from esc_gnn_b200.data import Data
    from esc_gnn_b200.distance import Distance
    from tests import batch_cases as BC
    fix = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'batch.npz'))
    for name, kw, has_attr, one_d, original in BC.DIST_CASES:
        inp = {k.split('/', 2)[2][3:]: torch.as_tensor(fix[k]) for k in fix.files if k.startswith('dist/%s/in_' % name)}
        assert ('edge_attr' in inp) == has_attr and ('original_pos' in inp) == original
        for dev in ('cpu', 'cuda'):
            d = Data(x=torch.ones(inp['pos'].size(0), 1), **{k: v.clone().to(dev) for k, v in inp.items()})
            o = Distance(**kw)(d)
            want = torch.as_tensor(fix['dist/%s/edge_attr' % name])
            assert o.edge_attr.shape == want.shape and o.edge_attr.dtype == want.dtype, name
            torch.testing.assert_close(o.edge_attr.cpu(), want, rtol=1e-6, atol=1e-6)
            if original:
                torch.testing.assert_close(o.original_edge_attr.cpu(), torch.as_tensor(fix['dist/%s/original_edge_attr' % name]),
                                           rtol=1e-6, atol=1e-6)


def test_ops_refuse_cpu_tensors():
    from esc_gnn_b200 import ops
    with pytest.raises(RuntimeError):
        ops.GraphIndex(torch.tensor([[0, 1], [1, 0]]), 2)


def _engine_for(variant, config, count, kw, use_graph, **extra):
    from esc_gnn_b200 import synth
    from esc_gnn_b200.engine import StaticTrainEngine
    from esc_gnn_b200.pipeline import RawBatch
    raw = RawBatch.synth(config, 100, count)
    model = build_product_model(variant, kw).cuda()
    sd = MU.det_state(model.state_dict(), seed=1234)
    model.load_state_dict({k: v.cuda() for k, v in sd.items()})
    model.train()
    fl = synth.ENCODER_FLAGS[config]
    eng = StaticTrainEngine(model, variant, fl, max_graphs=count, max_nodes_per_graph=max(64, raw.max_nodes),
                            max_edges_per_graph=max(256, raw.max_loop_edges),
                            nodes_cap=raw.num_nodes + 300, edges_cap=raw.src.numel() + 700, lr=1e-3, use_graph=use_graph,
                            **extra)
    return eng, model, raw


@pytest.mark.parametrize('name', ['zinc', 'count_h64'])
def test_engine_small_batch_after_large_batch_has_no_stale_rows(name):
    """Capacity rows past the batch's row count must stay inert: a small batch run right after a large one (whose rows filled
    the buffers further) gives the same loss and gradients as the same small batch on a fresh engine (lr = 0)."""
    from esc_gnn_b200.pipeline import RawBatch
    variant, config, count, kw = MU.MODEL_CASES[name]
    torch.backends.cuda.matmul.allow_tf32 = False
    raw0 = RawBatch.synth(config, 100, count)                       # _engine_for sizes the capacities from this batch
    raws = [RawBatch.synth(config, 500 + 31 * i, count) for i in range(10)]
    raws = [r for r in raws if r.num_nodes <= raw0.num_nodes + 300 and r.src.numel() <= raw0.src.numel() + 700]
    big, small = max(raws, key=lambda r: r.num_nodes), min(raws, key=lambda r: r.num_nodes)
    assert big.num_nodes >= small.num_nodes + 16 and big.src.numel() > small.src.numel()
    out = []
    for history in ([big, big, small], [small]):
        eng, _, raw0 = _engine_for(variant, config, count, kw, use_graph=False)
        assert big.num_nodes <= eng.c.caps['N'] and big.src.numel() <= eng.c.caps['E_in']
        eng.opt.hyper[0] = 0.0
        eng.opt._hyper_host = (0.0, 1.0); eng.opt.param_groups[0]['lr'] = 0.0
        for r in history:
            loss = float(eng.step(r).item())
        eng.check_errors()
        out.append((loss, eng.opt.grad.clone()))
    (la, ga), (lb, gb) = out
    assert abs(la - lb) <= 1e-6 * max(1.0, abs(lb))
    # two engines, same batch: what differs is the order of the fp32 vector reductions of the split-K weight gradients
    assert (ga - gb).abs().max().item() <= 2e-4 * gb.abs().max().item()       # float atomics reorder sums run to run (~3e-5); a stale row is O(1)


def test_engine_ordered_and_atomic_weight_gradients_agree():
    """atomic_wgrad=False (split-K partial tiles + ordered reduction) and the default (vector reductions into the zeroed
    gradient buffer) compute the same gradients."""
    variant, config, count, kw = MU.MODEL_CASES['zinc']
    torch.backends.cuda.matmul.allow_tf32 = False
    grads, losses = [], []
    for atomic, fused in ((False, False), (True, False), (True, True)):      # last: the one-launch readout tail as well
        eng, model, raw = _engine_for(variant, config, count, kw, use_graph=False, atomic_wgrad=atomic, fused_head=fused)
        eng.opt.hyper[0] = 0.0
        eng.opt._hyper_host = (0.0, 1.0); eng.opt.param_groups[0]['lr'] = 0.0
        losses.append(float(eng.step(raw).item()))
        grads.append(eng.opt.grad.clone())
    scale = grads[0].abs().max().item()
    for g, l in zip(grads[1:], losses[1:]):
        assert (grads[0] - g).abs().max().item() <= 2e-4 * scale
        assert abs(l - losses[0]) <= 1e-6 * max(1.0, abs(losses[0]))


@pytest.mark.parametrize('name', ['zinc', 'count_h64'])
def test_pipelined_engine_matches_sequential_engine(name):
    """pipeline=True (encoder of batch k overlapped with the training of batch k-1, staged batch set copied live at the
    start of the next step) produces the sequential engine's loss trajectory, one call late, over eager steps, the
    capture and graph replays, on batches of different sizes."""
    from esc_gnn_b200.pipeline import RawBatch
    variant, config, count, kw = MU.MODEL_CASES[name]
    torch.backends.cuda.matmul.allow_tf32 = False
    seq, _, raw0 = _engine_for(variant, config, count, kw, True)
    pip, _, _ = _engine_for(variant, config, count, kw, True, pipeline=True, encoder_ctas=8)      # encoder confined to 8 CTAs
    raws = [raw0] + [RawBatch.synth(config, 100 + 7 * i, count) for i in (1, 2)]
    raws = [r for r in raws if r.num_nodes <= seq.c.caps['N'] and r.src.numel() <= seq.c.caps['E_in']]
    assert len(raws) >= 2
    order = [raws[i % len(raws)] for i in range(7)]
    want = [float(seq.step(r).item()) for r in order]
    assert pip.step(order[0]) is None
    got = [float(pip.step(r).item()) for r in order[1:]] + [float(pip.drain().item())]
    seq.check_errors(); pip.check_errors()
    # Two SEQUENTIAL engines fed the same batches already differ by 1e-5 / 1e-4 / 1e-3 / 5e-3 after 2 / 3 / 5 / 7 steps
    # (float atomics in the embedding-table gradients reorder from run to run and Adam amplifies it, tools/debug_pipe.py):
    # the first steps are the tight check, the later ones only have to stay on the same trajectory.
    np.testing.assert_allclose(got[:3], want[:3], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(got[3:], want[3:], rtol=3e-2, atol=1e-5)


@pytest.mark.parametrize('pipeline', [False, True])
def test_step_read_returns_every_loss_one_call_late(pipeline):
    """engine.step_read: the host's copy of the loss comes from a pinned slot behind each step (no per-step device drain); the values
    are exactly the device losses of step(), one call later, and last_read() hands out the final one."""
    from esc_gnn_b200.pipeline import RawBatch
    variant, config, count, kw = MU.MODEL_CASES['zinc']
    torch.backends.cuda.matmul.allow_tf32 = False
    a, _, raw0 = _engine_for(variant, config, count, kw, True, pipeline=pipeline)
    b, _, _ = _engine_for(variant, config, count, kw, True, pipeline=pipeline)
    order = [raw0, RawBatch.synth(config, 107, count), raw0, raw0, RawBatch.synth(config, 107, count), raw0]
    order = [r for r in order if r.num_nodes <= a.c.caps['N'] and r.src.numel() <= a.c.caps['E_in']]
    want = []
    for r in order:
        l = a.step(r)
        want.append(None if l is None else float(l.item()))
    got = [b.step_read(r) for r in order] + [b.last_read()]
    assert got[0] is None                                   # nothing to hand out before the first step has run
    assert got[1:] == want or np.allclose([x for x in got[1:] if x is not None], [x for x in want if x is not None], rtol=1e-3)
    assert [x is None for x in got[1:]] == [x is None for x in want]


@pytest.mark.parametrize('name,use_graph', [('zinc', False), ('zinc', True), ('count_h256', True), ('count_h64', False),
                                            ('zinc_l2', True)])
def test_static_engine_train_steps_match_reference_fixture(name, use_graph):
    """Three full engine steps (encode -> collate -> fwd -> bwd -> Adam; step 3 is a CUDA-graph replay when use_graph)
    reproduce the loss trajectory and the post-training predictions of the reference's own class + torch Adam."""
    variant, config, count, kw = MU.MODEL_CASES[name]
    torch.backends.cuda.matmul.allow_tf32 = False
    eng, model, raw = _engine_for(variant, config, count, kw, use_graph)
    losses = [float(eng.step(raw).item()) for _ in range(3)]
    eng.check_errors()
    np.testing.assert_allclose(losses, FIX_M[name + '/adam_losses'], rtol=5e-3 if name != 'zinc_l2' else 2e-2, atol=1e-4)
    assert abs(losses[0] - FIX_M[name + '/loss'][0]) <= RTOL * abs(FIX_M[name + '/loss'][0]) + 1e-5
    model.eval()
    with torch.no_grad():
        pred = model(product_batch(config, 100, count)).cpu().numpy()
    want = FIX_M[name + '/pred_after_adam']
    # count variant: data.x is all ones, so x_embedding's first BatchNorm sees a zero-variance column; its eval-mode
    # output (const - running_mean) * rsqrt(running_var + eps) amplifies the rounding-level Adam steps of that layer
    # (same in the reference).  The training-mode losses above are the tight check there.
    tol = 1e-1 if variant == 'count' else 2e-2
    assert np.abs(pred - want).max() <= tol * max(np.abs(want).max(), 1.0)


@pytest.mark.parametrize('name', ['ogb', 'ogb_full'])
def test_static_engine_ogb_variant_matches_reference_fixture(name):
    """GNN(gin_eff) with virtual node through the engine (encode -> E1 on the bond columns -> forward -> BCE -> hand-written
    backward -> Adam): loss trajectory and post-training predictions of the reference's own class + torch Adam (drop_ratio 0)."""
    variant, config, count, kw = MU.MODEL_CASES[name]
    torch.backends.cuda.matmul.allow_tf32 = False
    eng, model, raw = _engine_for(variant, config, count, kw, True)
    losses = [float(eng.step(raw).item()) for _ in range(3)]
    eng.check_errors()
    assert abs(losses[0] - FIX_M[name + '/loss'][0]) <= RTOL * abs(FIX_M[name + '/loss'][0]) + 1e-5
    np.testing.assert_allclose(losses[:2], FIX_M[name + '/adam_losses'][:2], rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(losses[2], FIX_M[name + '/adam_losses'][2], rtol=5e-2, atol=1e-4)
    model.eval()
    with torch.no_grad():
        pred = model(product_batch(config, 100, count)).cpu().numpy()
    want = FIX_M[name + '/pred_after_adam']
    if name == 'ogb_full':
        # eval mode here mixes three steps of batch statistics over 8 graphs into the (random) initial running statistics of 12
        # virtual-node BatchNorms: the logits reach |40| and move by tens when a rounding-level gradient flips an Adam step.
        # The training-mode trajectory above, and gradients + running statistics against autograd (next test), are the checks.
        assert np.isfinite(pred).all() and np.sign(pred[np.abs(want) > 20]).tolist() == np.sign(want[np.abs(want) > 20]).tolist()
        return
    assert np.abs(pred - want).max() <= 5e-2 * max(np.abs(want).max(), 1.0)


@pytest.mark.parametrize('name', ['ogb', 'ogb_full'])
def test_static_engine_ogb_gradients_match_module_path(name):
    """Same batch, same weights, lr = 0: loss, every parameter gradient and every BatchNorm running statistic of the engine's
    hand-written OGB step equal autograd through the drop-in module."""
    variant, config, count, kw = MU.MODEL_CASES[name]
    torch.backends.cuda.matmul.allow_tf32 = False
    eng, model, raw = _engine_for(variant, config, count, kw, use_graph=False)
    eng.opt.hyper[0] = 0.0
    eng.opt._hyper_host = (0.0, 1.0); eng.opt.param_groups[0]['lr'] = 0.0
    loss_e = float(eng.step(raw).item())
    ref = build_product_model(variant, kw).cuda()
    sd = MU.det_state(ref.state_dict(), seed=1234)
    ref.load_state_dict({k: v.cuda() for k, v in sd.items()})
    ref.train()
    b = product_batch(config, 100, count)
    loss_m = MU.loss_fn(variant, ref(b), b.y)
    loss_m.backward()
    assert abs(loss_e - loss_m.item()) < 1e-5 * max(1.0, abs(loss_m.item())), (loss_e, loss_m.item())
    named_e = dict(model.named_parameters())
    bad = []
    for k, p in ref.named_parameters():
        if p.grad is None:
            continue
        ge = named_e[k].grad
        scale = max(p.grad.abs().max().item(), 1e-6)
        err = (ge - p.grad).abs().max().item()
        if err > 2e-2 * scale + 1e-6:
            bad.append((k, err, scale))
    bufs_e = dict(model.named_buffers())
    for k, v in ref.named_buffers():
        if k.endswith('running_mean') or k.endswith('running_var'):
            err = (bufs_e[k] - v).abs().max().item()
            if err > 1e-4 * max(v.abs().max().item(), 1.0):
                bad.append((k, err, 'running stat'))
    assert not bad, bad


def test_dropout_kernel_is_an_unbiased_reproducible_mask():
    import ctypes
    from esc_gnn_b200 import _lib
    L = _lib.lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rows, cap, C, p = 900, 1000, 300, 0.5
    x = torch.ones(cap, C, device='cuda')
    d_rows = torch.tensor([rows], dtype=torch.int32, device='cuda')
    step = torch.tensor([7], dtype=torch.int64, device='cuda')
    y1, y2, y3 = (torch.full((cap, C), 9.0, device='cuda') for _ in range(3))
    for y, s in ((y1, step), (y2, step), (y3, step + 1)):
        _lib.check(L.escgnn_dropout(P(x), C, p, 3, P(s), P(d_rows), cap, C, P(y), C, st), 'dropout')
    assert torch.equal(y1, y2) and not torch.equal(y1, y3)                 # same (salt, step) -> same mask; next step -> new mask
    assert float(y1[rows:].abs().max()) == 0.0
    kept = (y1[:rows] > 0).float().mean().item()
    assert abs(kept - 0.5) < 0.01 and set(y1[:rows].unique().tolist()) == {0.0, 2.0}
    assert abs(y1[:rows].mean().item() - 1.0) < 0.02                         # inverted scaling keeps the expectation


def test_static_engine_gradients_match_module_path():
    """Same batch, same weights: the engine's hand-written backward equals autograd through the module path."""
    variant, config, count, kw = MU.MODEL_CASES['zinc']
    torch.backends.cuda.matmul.allow_tf32 = False
    eng, model, raw = _engine_for(variant, config, count, kw, use_graph=False)
    eng.opt.hyper[0] = 0.0                    # lr = 0: parameters stay put, gradients stay in the flat buffer
    eng._hyper_host = None
    eng.opt._hyper_host = (0.0, 1.0); eng.opt.param_groups[0]['lr'] = 0.0
    loss_e = float(eng.step(raw).item())
    g_engine = eng.opt.grad.clone()
    ref = build_product_model(variant, kw).cuda()
    sd = MU.det_state(ref.state_dict(), seed=1234)
    ref.load_state_dict({k: v.cuda() for k, v in sd.items()})
    ref.train()
    b = product_batch(config, 100, count)
    loss_m = MU.loss_fn(variant, ref(b), b.y)
    loss_m.backward()
    assert abs(loss_e - loss_m.item()) < 1e-5 * max(1.0, abs(loss_m.item()))
    named_e = dict(model.named_parameters())
    bad = []
    for k, p in ref.named_parameters():
        ge = named_e[k].grad
        scale = max(p.grad.abs().max().item(), 1e-6)
        err = (ge - p.grad).abs().max().item()
        # both backward passes carry ~1e-3 (max-norm) fp32 conditioning noise in the early layers of this 5-layer BN/ELU
        # stack (measured against an fp64 run of the oracle: tools/debug_engine_grads.py); a wiring bug shows up as O(1)
        if err > 2e-2 * scale + 1e-6:
            bad.append((k, err, scale))
    assert not bad, bad


def test_static_engine_handles_varying_batches_under_one_graph():
    """One captured graph, different batches: each replay equals an eager run of the same step on the same batch."""
    from esc_gnn_b200.pipeline import RawBatch
    variant, config, count, kw = MU.MODEL_CASES['zinc']
    torch.backends.cuda.matmul.allow_tf32 = False
    eng_g, _, _ = _engine_for(variant, config, count, kw, use_graph=True)
    eng_e, _, _ = _engine_for(variant, config, count, kw, use_graph=False)
    for i in range(6):
        raw = RawBatch.synth(config, 2000 + 97 * i, count)
        lg = float(eng_g.step(raw).item()); le = float(eng_e.step(raw).item())
        # bag-embed backward atomics + Adam sign noise: two engines agree to ~1e-5 on the first steps and drift apart by up to
        # ~5e-4 after five (measured between two IDENTICAL sequential engines, tools/debug_pipe.py)
        assert abs(lg - le) <= (2e-4 if i < 2 else 3e-3) * max(1.0, abs(le)), (i, lg, le)
    # Parameters are NOT compared entry-wise: Adam turns rounding-level gradients (bag-embed backward uses float atomics)
    # into +-lr steps, so tiny-gradient entries legitimately differ between two runs.  The function they compute agrees:
    b = product_batch(config, 100, count)
    eng_g.model.eval(); eng_e.model.eval()
    with torch.no_grad():
        pg, pe = eng_g.model(b), eng_e.model(b)
    assert (pg - pe).abs().max().item() <= 5e-3 * max(1.0, pe.abs().max().item())


def test_bag_embed_backward_index_major_matches_atomic_version():
    import ctypes
    from esc_gnn_b200 import _lib, synth
    from esc_gnn_b200.transform import encode_batch
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    src, dst, eptr, nptr = synth.make_batch_arrays(2, 4000, 64)
    r = encode_batch(torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda(), torch.as_tensor(eptr), torch.as_tensor(nptr),
                     3, True, False, expand=True)
    E, H = r.num_edges, 256
    g = torch.randn(E + 50, H, device='cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    d_count = torch.tensor([E], dtype=torch.int32, device='cuda')
    dW = torch.zeros(1800, H, device='cuda')
    work = torch.zeros(3 * 1800 + 8, dtype=torch.int32, device='cuda')
    cap = r.nnz + 1000
    se, sc = torch.zeros(cap, dtype=torch.int32, device='cuda'), torch.zeros(cap, device='cuda')
    _lib.check(_lib.lib().escgnn_bag_embed_bwd_sorted(P(g), H, P(r.rec), P(r.rec_off), P(r.rec_nnz), E + 50, cap, P(dW), P(work), P(se),
                                                      P(sc), P(d_count), st), 'bag_embed_bwd_sorted')
    ref = torch.zeros(1800, H, device='cuda', dtype=torch.float64)
    ref.index_add_(0, r.pos_index, g[r.pos_batch].double() * r.pos_enc.view(-1, 1).double())
    # fp32 accumulation (float atomics across chunks, order varies) of terms far larger than their sum: the absolute error
    # scales with sum |terms| (entries reach ~1e3 here; 1.4e-4 observed on one entry in one of six runs)
    mag = torch.zeros(1800, H, device='cuda', dtype=torch.float64)
    mag.index_add_(0, r.pos_index, g[r.pos_batch].double().abs() * r.pos_enc.view(-1, 1).double())
    assert ((dW.double() - ref).abs() <= 1e-5 * ref.abs() + 2e-6 * mag + 1e-6).all()


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json shapes (batch 128 / 256 / 32 / 32, 5-6 layers, hidden 256 / 300) and the fp64 gradient yardstick
def _fp64_truth(name):
    """fp64 run of the oracle restatement (plain torch, on the GPU for speed) -- first pinned to the fp64 run of the reference's
    own class (fixture `/grad64_digest`, 1e-9), then used as the truth both backward passes are measured against."""
    variant, config, count, kw = MU.MODEL_CASES[name]
    m = MU.build_oracle_model(variant, kw).double()
    sd = MU.det_state(m.state_dict(), seed=1234)
    m.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()})
    m = m.cuda().train()
    b = MU.to_double(MU.ref_batch(config, 100, count))
    for k, v in list(b.__dict__.items()):
        if torch.is_tensor(v):
            setattr(b, k, v.cuda())
    loss = MU.loss_fn(variant, m(b), b.y) if variant != 'ogb' else None
    if variant == 'ogb':
        pred = m(b)
        y = b.y.view(pred.shape)
        loss = torch.nn.BCEWithLogitsLoss()(pred[y == y], y[y == y])
    loss.backward()
    g64 = {k: p.grad.detach() for k, p in m.named_parameters() if p.grad is not None}
    assert abs(loss.item() - FIX_M[name + '/loss64'][0]) <= 1e-9 * max(1.0, abs(FIX_M[name + '/loss64'][0]))
    for k, want in zip(FIX_M[name + '/grad_keys'], FIX_M[name + '/grad64_digest']):
        if variant == 'count' and str(k).startswith('x_embedding.'):
            continue                  # all-ones input: mathematically zero gradients (norms of 1e-12 that depend on the summation order)
        got = MU.grad_digest(g64[str(k)])
        assert abs(got[3] - want[3]) <= 1e-8 * max(want[3], 1e-12) + 1e-13, (str(k), got[3], want[3])     # ||g64||_2
    return g64


def _check_against_fp64(name, grads, g64, factor):
    """||g - g64||_2 <= factor * ||g32_reference - g64||_2 + floor, per parameter tensor.  floor: 2e-6 of the tensor's norm
    (fp32 resolution of the sum itself) -- it only matters where the reference's own error is ~0, i.e. gradients that are
    mathematically zero (a Linear bias feeding BatchNorm), which the engine leaves at exactly 0."""
    variant = MU.MODEL_CASES[name][0]
    worst, bad, strict = 0.0, [], []
    gmax = max(float(v.norm()) for v in g64.values())
    from tests.test_model_oracle_cpu import reference_worst_relative_error
    case_rel = 10.0 * reference_worst_relative_error(name, variant)
    for k, ref_err in zip(FIX_M[name + '/grad_keys'], FIX_M[name + '/grad_err32']):
        k = str(k)
        if variant == 'count' and k.startswith('x_embedding.'):
            continue                  # all-ones input: zero-variance BatchNorm, gradients are rounding noise times rsqrt(eps) (see run_case)
        err = float((grads[k].double() - g64[k]).norm())
        # ... or, where the reference happened to land much closer to the truth than it typically does, 10x the relative error
        # of the reference's own worst tensor in this case.  The product's floor is the tensor core: tcgen05.mma accumulates
        # with truncation (measured: -2.6e-6 mean signed relative error at K = 256 against +3e-10 for an FFMA GEMM,
        # tools/bench_linear_bn.py), which the BatchNorm stacks amplify like any other fp32 rounding.
        strict_bound = factor * float(ref_err) + 2e-6 * float(g64[k].norm()) + 1e-9 * gmax
        strict.append(err / max(strict_bound, 1e-30))
        bound = max(strict_bound, case_rel * float(g64[k].norm()) + 1e-9 * gmax)
        worst = max(worst, err / max(bound, 1e-30))
        if err > bound:
            bad.append((k, err, float(ref_err), float(g64[k].norm())))
    assert not bad, (name, worst, bad[:6])
    # The STRICT per-tensor bound (3x the reference's own fp32 error on that tensor) holds for the typical tensor -- the median
    # ratio is 0.2 .. 0.5, i.e. the product is usually CLOSER to the fp64 truth than the reference's CPU run.  It cannot hold for
    # every tensor of every case: a ReLU whose pre-activation is within rounding of zero takes the other branch under any fp32
    # perturbation (tools/debug_relu_flips.py: 5 of the 12.7 M GINE message entries of the 256-graph ZINC batch do, with
    # |pre-activation| < 1.6e-6), and each crossing shifts the gradients downstream of it by ~1e-4 of their norm -- which is what
    # the looser case-level bound above absorbs.
    # One crossing early in a ReLU network (count / ogb variants: BatchNorm -> ReLU on every layer) moves EVERY gradient, so
    # whether the median meets the strict bound there is decided by which rounding perturbation a build happens to have
    # (tools/debug_grad_error.py count_cfg3: the same product kernels give median 0.31 / worst 0.97 with the two-kernel BatchNorm
    # and median 5.3 / worst 1646 with the one-launch BatchNorm -- and so does plain torch: 0.50 / 230 with cuBLAS + torch BN).
    # The ELU network of the headline config has kinks only inside the GINE messages: the median is asserted there.
    strict.sort()
    if variant == 'zinc':
        assert strict[len(strict) // 2] <= 1.0, (name, 'median strict ratio', strict[len(strict) // 2])
    return worst


@pytest.mark.parametrize('name', list(MU.FP64_CASES))
def test_gradients_within_3x_of_reference_fp32_error(name):
    """Drop-in module (autograd over the sm_100a kernels) AND the engine's hand-written backward, at the reference's own batch
    sizes: per tensor, the distance to the fp64 gradient is at most 3x the distance of the REFERENCE's fp32 gradient to it."""
    variant, config, count, kw = MU.MODEL_CASES[name]
    torch.backends.cuda.matmul.allow_tf32 = False
    g64 = _fp64_truth(name)
    # module path
    ref = build_product_model(variant, kw).cuda()
    sd = MU.det_state(ref.state_dict(), seed=1234)
    ref.load_state_dict({k: v.cuda() for k, v in sd.items()})
    ref.train()
    b = product_batch(config, 100, count)
    MU.loss_fn(variant, ref(b), b.y).backward()
    w_mod = _check_against_fp64(name, {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}, g64, 3.0)
    # engine path (lr = 0 keeps the gradients in the flat buffer)
    eng, model, raw = _engine_for(variant, config, count, kw, use_graph=False)
    eng.opt.hyper[0] = 0.0
    eng.opt._hyper_host = (0.0, 1.0); eng.opt.param_groups[0]['lr'] = 0.0
    loss_e = float(eng.step(raw).item())
    eng.check_errors()
    assert abs(loss_e - FIX_M[name + '/loss64'][0]) <= RTOL * abs(FIX_M[name + '/loss64'][0]) + 1e-5
    w_eng = _check_against_fp64(name, {k: p.grad for k, p in model.named_parameters()}, g64, 3.0)
    print('%s: worst err/bound module %.2f engine %.2f' % (name, w_mod, w_eng))


@pytest.mark.parametrize('name', ['count_cfg1', 'zinc_cfg2', 'count_cfg3'])
def test_static_engine_baseline_shapes_match_reference_fixture(name):
    """Three engine steps at BASELINE.json's batch sizes (cfg 1: 128 graphs h=3, cfg 2: 256 graphs, cfg 3: 32 graphs h=4) against
    the loss trajectory of the reference's own class + torch Adam."""
    variant, config, count, kw = MU.MODEL_CASES[name]
    torch.backends.cuda.matmul.allow_tf32 = False
    eng, model, raw = _engine_for(variant, config, count, kw, True)
    losses = [float(eng.step(raw).item()) for _ in range(3)]
    eng.check_errors()
    assert abs(losses[0] - FIX_M[name + '/loss'][0]) <= RTOL * abs(FIX_M[name + '/loss'][0]) + 1e-5
    np.testing.assert_allclose(losses[:2], FIX_M[name + '/adam_losses'][:2], rtol=5e-3, atol=1e-4)
    # third loss: two Adam steps turn rounding-level gradient differences into +-lr parameter steps (cfg 3: 32 graphs, the loss
    # moves 0.62 -> 1.24 -> 1.06 in these steps)
    np.testing.assert_allclose(losses[2], FIX_M[name + '/adam_losses'][2], rtol=2e-2, atol=1e-4)


def test_static_engine_ogb_baseline_shape_matches_reference_fixture():
    name = 'ogb_cfg4'
    variant, config, count, kw = MU.MODEL_CASES[name]
    torch.backends.cuda.matmul.allow_tf32 = False
    eng, model, raw = _engine_for(variant, config, count, kw, True)
    losses = [float(eng.step(raw).item()) for _ in range(3)]
    eng.check_errors()
    assert abs(losses[0] - FIX_M[name + '/loss'][0]) <= RTOL * abs(FIX_M[name + '/loss'][0]) + 1e-5
    np.testing.assert_allclose(losses[:2], FIX_M[name + '/adam_losses'][:2], rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(losses[2], FIX_M[name + '/adam_losses'][2], rtol=5e-2, atol=1e-4)


def test_zinc_engine_with_self_loops_rewrites_bond_types():
    """flags['self_loop'] on the ZINC variant (run_zinc.py --self_loop): bond types go through E1 with the edge list (loop rows
    dropped, N rows of 1 appended); engine loss and gradients equal the drop-in module fed the reference-contract batch."""
    from esc_gnn_b200 import synth
    from esc_gnn_b200.engine import StaticTrainEngine
    from esc_gnn_b200.pipeline import RawBatch
    torch.backends.cuda.matmul.allow_tf32 = False
    kw, count = dict(num_layers=3), 24
    raw = RawBatch.synth(2, 100, count)
    fl = dict(synth.ENCODER_FLAGS[2], self_loop=True)
    for pipeline in (False, True):
        model = build_product_model('zinc', kw).cuda()
        sd = MU.det_state(model.state_dict(), seed=1234)
        model.load_state_dict({k: v.cuda() for k, v in sd.items()})
        model.train()
        eng = StaticTrainEngine(model, 'zinc', fl, max_graphs=count, max_nodes_per_graph=64, max_edges_per_graph=256,
                                nodes_cap=raw.num_nodes + 100, edges_cap=raw.src.numel() + 200, lr=0.0, use_graph=False,
                                pipeline=pipeline)
        eng.opt.hyper[0] = 0.0
        eng.opt._hyper_host = (0.0, 1.0); eng.opt.param_groups[0]['lr'] = 0.0
        if pipeline:
            assert eng.step(raw) is None
            loss_e = float(eng.drain().item())
        else:
            loss_e = float(eng.step(raw).item())
        eng.check_errors()
        ref = build_product_model('zinc', kw).cuda()
        ref.load_state_dict({k: v.cuda() for k, v in sd.items()})
        ref.train()
        from esc_gnn_b200.batch import Batch
        b = Batch.from_data_list([to_data(g) for g in MU.graph_dicts(2, 100, count, self_loop=True)]).to('cuda')
        assert int((b.edge_attr == 1).sum()) >= b.x.size(0)          # the appended loop rows carry bond type 1
        loss_m = MU.loss_fn('zinc', ref(b), b.y)
        loss_m.backward()
        assert abs(loss_e - loss_m.item()) <= 1e-5 * max(1.0, abs(loss_m.item())), (pipeline, loss_e, loss_m.item())
        named_e = dict(model.named_parameters())
        for k, p in ref.named_parameters():
            err = (named_e[k].grad - p.grad).abs().max().item()
            assert err <= 2e-2 * max(p.grad.abs().max().item(), 1e-6) + 1e-6, (k, err)


def test_engine_error_counters_are_sticky_and_overflow_is_contained():
    """A batch that exceeds the record capacity must not corrupt memory and must still be reported by check_errors() after LATER
    clean batches (the per-call counters are zeroed every step, the sticky slots are not)."""
    from esc_gnn_b200 import synth
    from esc_gnn_b200.engine import StaticTrainEngine
    from esc_gnn_b200.pipeline import RawBatch
    count = 16
    seeds = [100 + 50 * i for i in range(10)]
    nnz = {s: int(MU.ref_batch(2, s, count).pos_enc.numel()) for s in seeds}
    s_small, s_big = min(nnz, key=nnz.get), max(nnz, key=nnz.get)
    small, big = RawBatch.synth(2, s_small, count), RawBatch.synth(2, s_big, count)
    edges_cap = max(small.src.numel(), big.src.numel()) + 8
    rpe = -(-nnz[s_small] // edges_cap)
    assert edges_cap * rpe < nnz[s_big], 'seeds do not separate: %r' % (nnz, )
    model = build_product_model('zinc', dict(num_layers=2)).cuda().train()
    eng = StaticTrainEngine(model, 'zinc', synth.ENCODER_FLAGS[2], max_graphs=count, max_nodes_per_graph=64, max_edges_per_graph=256,
                            nodes_cap=max(small.num_nodes, big.num_nodes) + 8, edges_cap=edges_cap, lr=1e-3, use_graph=False,
                            records_per_edge=rpe)
    loss = eng.step(big)                                             # overflows
    torch.cuda.synchronize()
    assert bool(torch.isfinite(loss).all())
    nnz_edges = eng.rec_nnz[:int(eng.c.dims[1])]
    assert int(nnz_edges.sum()) <= eng.rec.numel() and int((nnz_edges == 0).sum()) > 0     # edges without room own no records
    assert int((eng.rec_off[:int(eng.c.dims[1])] + nnz_edges).max()) <= eng.rec.numel()
    eng.step(small)                                                  # clean batch afterwards: per-call counters are reset ...
    with pytest.raises(RuntimeError, match='record capacity'):
        eng.check_errors()                                           # ... but the overflow is still reported
    eng.step(small)
    eng.check_errors()                                               # cleared by the read; clean since


@pytest.mark.parametrize('name', ['zinc', 'count_h64', 'ogb'])
def test_engine_trains_on_a_partial_last_batch(name):
    """The last batch of an epoch has fewer graphs than the engine's capacity (the reference trains on it): same loss and
    gradients as an engine built for exactly that many graphs."""
    from esc_gnn_b200 import synth
    from esc_gnn_b200.engine import StaticTrainEngine
    from esc_gnn_b200.pipeline import RawBatch
    variant, config, count, kw = MU.MODEL_CASES[name]
    torch.backends.cuda.matmul.allow_tf32 = False
    part = max(2, count * 2 // 3)
    full, raw = RawBatch.synth(config, 100, count), RawBatch.synth(config, 100, part)
    out = []
    for G in (count, part):
        model = build_product_model(variant, kw).cuda()
        sd = MU.det_state(model.state_dict(), seed=1234)
        model.load_state_dict({k: v.cuda() for k, v in sd.items()})
        model.train()
        eng = StaticTrainEngine(model, variant, synth.ENCODER_FLAGS[config], max_graphs=G, max_nodes_per_graph=max(64, full.max_nodes),
                                max_edges_per_graph=max(256, full.max_loop_edges), nodes_cap=full.num_nodes + 300,
                                edges_cap=full.src.numel() + 700, lr=1e-3, use_graph=False)
        eng.opt.hyper[0] = 0.0
        eng.opt._hyper_host = (0.0, 1.0); eng.opt.param_groups[0]['lr'] = 0.0
        if G == count:
            eng.step(full)                       # a full batch first: its rows / graphs must not leak into the partial one
        loss = float(eng.step(raw).item())
        eng.check_errors()
        assert int(eng.c.dims[2]) == part
        out.append((loss, eng.opt.grad.clone()))
    (la, ga), (lb, gb) = out
    assert abs(la - lb) <= 1e-6 * max(1.0, abs(lb)), (la, lb)
    # two engines, same batch: what differs is the order of the fp32 vector reductions of the split-K weight gradients
    assert (ga - gb).abs().max().item() <= 2e-4 * gb.abs().max().item()


@pytest.mark.parametrize('name', ['zinc_cfg2', 'count_cfg1', 'zinc'])
def test_engine_fused_linear_bn_matches_separate_launches(name):
    """fuse_bn=True (Linear + BatchNorm + activation as one launch each way: GEMM epilogues behind a grid barrier) against the same
    engine with GEMM and BatchNorm as separate launches: loss, every gradient, running statistics -- and the graph replay of the
    fused tape over three Adam steps against the reference fixture."""
    variant, config, count, kw = MU.MODEL_CASES[name]
    torch.backends.cuda.matmul.allow_tf32 = False
    res = []
    for fuse in (True, False):
        eng, model, raw = _engine_for(variant, config, count, kw, use_graph=False, fuse_bn=fuse)
        eng.opt.hyper[0] = 0.0
        eng.opt._hyper_host = (0.0, 1.0); eng.opt.param_groups[0]['lr'] = 0.0
        loss = float(eng.step(raw).item())
        eng.check_errors()
        res.append((loss, {k: p.grad.clone() for k, p in model.named_parameters()},
                    {k: v.clone() for k, v in model.state_dict().items() if 'running' in k},
                    sum(1 for f in eng.fwd)))
    (lf, gf, rf, nf), (lu, gu, ru, nu) = res
    assert nf < nu                                         # the fused tape really is shorter
    assert abs(lf - lu) <= 2e-6 * max(1.0, abs(lu)), (lf, lu)
    gmax = max(float(v.abs().max()) for v in gu.values())
    for k in gu:
        if variant == 'count' and k.startswith('x_embedding.'):
            continue
        # eps scalars: sums of N * C cancelling products; everything else: two summation orders of the BatchNorm backward sums,
        # amplified through the layers below (the bag-embed gradient is the end of the chain)
        tol = 5e-3 if gu[k].numel() == 1 else 3e-3
        assert float((gf[k] - gu[k]).abs().max()) <= tol * float(gu[k].abs().max()) + 1e-7 * gmax, k
    for k in ru:
        if variant == 'count' and k.startswith('x_embedding.6.'):
            continue
        torch.testing.assert_close(rf[k], ru[k], rtol=2e-5, atol=2e-6)
    eng, model, raw = _engine_for(variant, config, count, kw, use_graph=True, fuse_bn=True)
    losses = [float(eng.step(raw).item()) for _ in range(3)]
    eng.check_errors()
    assert abs(losses[0] - FIX_M[name + '/loss'][0]) <= RTOL * abs(FIX_M[name + '/loss'][0]) + 1e-5
    np.testing.assert_allclose(losses, FIX_M[name + '/adam_losses'], rtol=5e-3 if name != 'zinc' else 2e-2, atol=1e-4)
