"""Engine throughput on the other BASELINE.json config shapes (not the judged bench line): count_cycle-shaped batches of 128
(h=3) and count_graphlet-shaped batches of 32 (h=4) through the count variant of NestedGIN_eff, ZINC at several batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esc_gnn_b200 import synth, zinc_model, graphcount_model
from esc_gnn_b200.engine import StaticTrainEngine
from esc_gnn_b200.pipeline import RawBatch


def run(tag, variant, config, batch, layers=5, hidden=256, steps=200, encoder_ctas=74):
    fl = synth.ENCODER_FLAGS[config]
    pool = [RawBatch.synth(config, 7000 + i * batch, batch).cuda(non_blocking=False) for i in range(6)]
    ncap = int(max(b.num_nodes for b in pool) * 1.04) + 64
    ecap = int(max(b.src.numel() for b in pool) * 1.04) + 128
    mx_n = max(b.max_nodes for b in pool)
    mx_e = max((b.max_loop_edges if fl['self_loop'] else b.max_in_edges) for b in pool)
    torch.manual_seed(0)
    if variant == 'zinc':
        model = zinc_model.NestedGIN_eff(None, layers).cuda()
    else:
        model = graphcount_model.NestedGIN_eff(None, layers, hidden, use_rd=True, graph_pred=False, dropout=0, edge_nest=True,
                                               use_cycle=True).cuda()
    model.train()
    eng = StaticTrainEngine(model, variant, fl, max_graphs=batch, max_nodes_per_graph=mx_n, max_edges_per_graph=mx_e, nodes_cap=ncap,
                            edges_cap=ecap, lr=1e-3, pipeline=True, encoder_ctas=encoder_ctas)
    losses = []
    for i in range(10):
        l = eng.step(pool[i % 6])
        if l is not None:
            losses.append(float(l.item()))
    eng.check_errors()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        eng.step(pool[i % 6])
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    eng.check_errors()
    d = eng.c.dims.cpu().tolist()
    if os.environ.get('ESC_PROFILE'):
        km, _ = eng.profile(pool[0], reps=5)
        print('   ', {k: round(v, 4) for k, v in sorted(km.items(), key=lambda kv: -kv[1])[:14]})
    print('%-48s batch %5d  N %6d E %7d nnz %8d  %.3f ms/step  %9.0f graphs/s   loss %.4f -> %.4f' % (
        tag, batch, d[0], d[1], d[3], ms, batch / ms * 1e3, losses[0], losses[-1]), flush=True)


for cap in (74, 0):
    run('cfg1 count_cycle-shaped h=3 (encoder cap %d)' % cap, 'count', 1, 128, encoder_ctas=cap)
    run('cfg3 count_graphlet-shaped h=4 (cap %d)' % cap, 'count', 3, 32, encoder_ctas=cap)
for bsz in (64, 256, 1024):
    run('cfg2 ZINC-shaped h=3', 'zinc', 2, bsz)


def run_module_path(tag, config, batch, steps=60):
    """Config 4 (ogbg-molhiv-shaped, h=4): GNN(gin_eff, virtual node) through the drop-in module path (torch autograd around the
    sm_100a kernels) -- the static engine has no OGB variant yet."""
    from esc_gnn_b200 import ogb_model
    from esc_gnn_b200.pipeline import TrainPipeline
    from tests.model_util import loss_fn
    fl = synth.ENCODER_FLAGS[config]
    pool = [RawBatch.synth(config, 9000 + i * batch, batch).cuda(non_blocking=False) for i in range(6)]
    torch.manual_seed(0)
    model = ogb_model.GNN('ogbg-molhiv', 1, num_layer=6, emb_dim=300, gnn_type='gin_eff', virtual_node=True, residual=False,
                          drop_ratio=0.5).cuda()
    model.train()
    pipe = TrainPipeline(model, lambda p, y: loss_fn('ogb', p, y), fl['h'], fl['use_rd'], fl['self_loop'], lr=1e-3)
    for i in range(5):
        l0 = float(pipe.step_device(pool[i % 6]).item())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        l = pipe.step_device(pool[i % 6])
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    print('%-48s batch %5d  %.3f ms/step  %9.0f graphs/s   loss %.4f -> %.4f (module path)' % (tag, batch, ms, batch / ms * 1e3, l0,
                                                                                              float(l.item())), flush=True)


run_module_path('cfg4 ogbg-molhiv-shaped h=4, 6 layers emb 300', 4, 32)


def run_ogb_engine(tag, batch, steps=200, drop_ratio=0.5):
    """Config 4 through the static engine (GNN gin_eff, 6 layers, emb 300, virtual node, dropout)."""
    from esc_gnn_b200 import ogb_model
    config = 4
    fl = synth.ENCODER_FLAGS[config]
    pool = [RawBatch.synth(config, 9000 + i * batch, batch).cuda(non_blocking=False) for i in range(6)]
    ncap = int(max(b.num_nodes for b in pool) * 1.04) + 64
    ecap = int(max(b.src.numel() for b in pool) * 1.04) + 128
    torch.manual_seed(0)
    model = ogb_model.GNN('ogbg-molhiv', 1, num_layer=6, emb_dim=300, gnn_type='gin_eff', virtual_node=True, residual=False,
                          drop_ratio=drop_ratio).cuda()
    model.train()
    eng = StaticTrainEngine(model, 'ogb', fl, max_graphs=batch, max_nodes_per_graph=max(b.max_nodes for b in pool),
                            max_edges_per_graph=max(b.max_loop_edges for b in pool), nodes_cap=ncap, edges_cap=ecap, lr=1e-3,
                            pipeline=True)
    losses = []
    for i in range(10):
        l = eng.step(pool[i % 6])
        if l is not None:
            losses.append(float(l.item()))
    eng.check_errors()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        l = eng.step(pool[i % 6])
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    eng.check_errors()
    print('%-48s batch %5d  %.3f ms/step  %9.0f graphs/s   loss %.4f -> %.4f (engine, dropout %.1f)' % (
        tag, batch, ms, batch / ms * 1e3, losses[0], float(l.item()), drop_ratio), flush=True)
    if os.environ.get('ESC_PROFILE'):
        km, _ = eng.profile(pool[0], reps=5)
        print('   ', {k: round(v, 4) for k, v in sorted(km.items(), key=lambda kv: -kv[1])[:16]})


run_ogb_engine('cfg4 ogbg-molhiv-shaped h=4, 6 layers emb 300', 32)
run_ogb_engine('cfg4 ogbg-molhiv-shaped h=4, 6 layers emb 300', 256)
