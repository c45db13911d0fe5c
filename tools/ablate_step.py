"""How much of the step time each kernel family is responsible for: time the pipelined engine with that family's entry
points replaced by no-ops (results are garbage, timing only).  usage: python tools/ablate_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esc_gnn_b200 import _lib, synth, zinc_model
from esc_gnn_b200.engine import StaticTrainEngine
from esc_gnn_b200.pipeline import RawBatch

FAMILIES = {
    'none': [],
    'bn': ['escgnn_bn_act_fwd', 'escgnn_bn_act_bwd'],
    'gemm_all': ['escgnn_gemm_tf32x3_bounded'],
    'gine': ['escgnn_gine_aggregate_fwd_ld', 'escgnn_gine_aggregate_bwd_ld'],
    'bag': ['escgnn_bag_embed_fwd', 'escgnn_bag_index_build', 'escgnn_bag_embed_bwd_indexed'],
    'encode_rd': ['escgnn_encode_rd'],
    'gemm_fwd_dgrad(M-bound)': None,
    'head': ['escgnn_head_bn_linear_l1'],
    'colsum+reduce_sum': ['escgnn_colsum', 'escgnn_reduce_sum'],
    'embedding+pool': ['escgnn_embedding_fwd', 'escgnn_embedding_bwd', 'escgnn_segment_pool_fwd', 'escgnn_segment_pool_bwd'],
}
L = _lib.lib()
orig = {}
pool = [RawBatch.synth(2, i * 256, 256).cuda(non_blocking=False) for i in range(6)]
ncap = int(max(b.num_nodes for b in pool) * 1.04) + 64
ecap = int(max(b.src.numel() for b in pool) * 1.04) + 128
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
base = None
for fam, names in FAMILIES.items():
    if names is None:
        continue
    for n in names:
        orig[n] = getattr(L, n)
        setattr(L, n, (lambda *a: 0))
    torch.manual_seed(0)
    model = zinc_model.NestedGIN_eff(None, 5).cuda(); model.train()
    eng = StaticTrainEngine(model, 'zinc', synth.ENCODER_FLAGS[2], max_graphs=256, max_nodes_per_graph=40, max_edges_per_graph=96,
                            nodes_cap=ncap, edges_cap=ecap, lr=1e-3, pipeline=True, encoder_ctas=74)
    for i in range(8):
        eng.step(pool[i % 6])
    torch.cuda.synchronize()
    tot = 0.0
    K = 200
    evs = []
    for i in range(K):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.step(pool[i % 6]); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / K
    if base is None:
        base = ms
    print('%-20s %.4f ms/step   (%.4f less than the full step)' % (fam, ms, base - ms), flush=True)
    for n in names:
        setattr(L, n, orig[n])
    del eng, model
