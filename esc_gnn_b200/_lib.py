"""ctypes binding of libescgnn_b200.so (C-ABI in include/escgnn_b200.h).

There is NO CPU fallback: importing this module without the built library, or calling a device entry point
without a CUDA device, raises.  Build with `python -m esc_gnn_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libescgnn_b200.so')

ERR_BITS = {1: 'sub_degree >= 200 (reference: F.one_hot(sub_degree, 200) raises)',
            2: 'node id outside [0, num_nodes)',
            4: 'resistance-distance bin outside the supported range / singular system',
            8: 'use_rd on a non-symmetric edge multiset is not supported'}
NUM_COUNTERS = 8
CTR_PER_CALL, CTR_STICKY_ERROR, CTR_MAX_NNZ = 4, 4, 5
RD_SLOTS = 12
REC_IDX_BITS = 11

_lib = None
_vp, _i64, _i32, _u32p = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)
_i64p = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes); every symbol include/escgnn_b200.h declares
SIGNATURES = {
    'escgnn_version': (_i32, []),
    'escgnn_rewrite_self_loops': (_i32, [_vp] * 4 + [_i64] + [_vp] * 5),
    'escgnn_encode': (_i32, [_vp] * 4 + [_i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    'escgnn_encode_subset': (_i32, [_vp] * 4 + [_i64, _vp, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    'escgnn_encode_scratch_bytes': (_i64, [_i64, _i64, _i32]),
    'escgnn_encode_rd_scratch_bytes': (_i64, [_i64, _i64, _i32]),
    'escgnn_encode_rd': (_i32, [_vp] * 4 + [_i64, _i32, _vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    'escgnn_exclusive_scan_i32': (_i32, [_vp, _i64, _vp, _vp, _vp]),
    'escgnn_expand_records': (_i32, [_vp] * 6 + [_i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    'escgnn_ctx_create': (_vp, [_i32]),
    'escgnn_ctx_destroy': (None, [_vp]),
    'escgnn_encode_host_run': (_i32, [_vp] * 5 + [_i64, _i32, _i32, _i32, _i32, _i64p, _i64p, _u32p]),
    'escgnn_encode_host_fetch': (_i32, [_vp] * 7),
    'escgnn_encode_host_submit': (_i32, [_vp, _i32] + [_vp] * 4 + [_i64, _i32, _i32, _i32]),
    'escgnn_encode_host_wait': (_i32, [_vp, _i32, _i64p, _i64p, _u32p] + [ctypes.POINTER(ctypes.c_void_p)] * 6),
    'escgnn_expand_records_host': (_i32, [_vp] * 4 + [_i64, _i32, _vp, _vp, _vp, _i32]),
    'escgnn_encode_host_device_results': (_i32, [_vp] * 6),
    'escgnn_csr_build': (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'escgnn_sorted_ids_to_ptr': (_i32, [_vp, _i64, _i64, _vp, _vp, _vp]),
    'escgnn_bag_embed_fwd': (_i32, [_vp, _i32] + [_vp] * 6 + [_i64, _vp, _vp, _vp]),
    'escgnn_bag_embed_bwd': (_i32, [_vp, _i32] + [_vp] * 6 + [_i64, _vp, _vp, _vp]),
    'escgnn_bag_index_build': (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    'escgnn_bag_embed_bwd_indexed': (_i32, [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    'escgnn_dropout': (_i32, [_vp, _i32, ctypes.c_float, ctypes.c_uint32, _vp, _vp, _i32, _i32, _vp, _i32, _vp]),
    'escgnn_add_segment_rows': (_i32, [_vp, _i32, _vp, _i32, _vp, _i64, _i32, _vp, _i32, _vp]),
    'escgnn_rewrite_edge_attr': (_i32, [_vp] * 4 + [_i64, _vp, _vp, _i32, _i64, _vp, _vp]),
    'escgnn_embedding_bwd_small': (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    'escgnn_reduce_sum': (_i32, [_vp, _i64, _vp, _i32, _vp]),
    'escgnn_zero_tail_rows': (_i32, [_vp, _i32, _i32, _vp, _i64, _vp]),
    'escgnn_bag_embed_bwd_sorted': (_i32, [_vp, _i32, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'escgnn_gine_aggregate_fwd': (_i32, [_vp] * 6 + [_i64, _i32, _vp, _vp, _vp]),
    'escgnn_gine_aggregate_bwd': (_i32, [_vp] * 7 + [_i64, _i32] + [_vp] * 6),
    'escgnn_gine_aggregate_fwd_ld': (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp]),
    'escgnn_gine_aggregate_bwd_ld': (_i32, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    'escgnn_segment_pool_fwd': (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    'escgnn_segment_pool_bwd': (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    'escgnn_collate_edges': (_i32, [_vp] * 4 + [_i64, _vp, _vp, _vp, _vp]),
    'escgnn_ptr_to_ids': (_i32, [_vp, _i64, _i64, _vp, _vp, _vp]),
    'escgnn_dense_tile_rows': (_i32, []),
    'escgnn_head_bn_linear_l1': (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    'escgnn_set_pdl': (_i32, [_i32]),
    'escgnn_set_encoder_grid_cap': (_i32, [_i32]),
    'escgnn_set_rd_peel': (_i32, [_i32]),
    'escgnn_set_rd_fast': (_i32, [_i32]),
    'escgnn_set_cluster_bn': (_i32, [_i32]),
    'escgnn_dense_partial_floats': (_i64, [_i32, _i32]),
    'escgnn_bn_act_fwd': (_i32, [_vp, _i32] + [_vp] * 7 + [_i32, ctypes.c_float, ctypes.c_float, _i32, _vp, _i32, _i32, _vp, _i32, _vp]),
    'escgnn_bn_act_bwd': (_i32, [_vp, _i32, _vp, _i32, _vp, _i32] + [_vp] * 4 + [_i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp]),
    'escgnn_act_fwd': (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp]),
    'escgnn_act_bwd': (_i32, [_vp, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp]),
    'escgnn_colsum': (_i32, [_vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    'escgnn_embedding_fwd': (_i32, [_vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp, _i32, _vp]),
    'escgnn_embedding_bwd': (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    'escgnn_loss_fwd_bwd': (_i32, [_vp, _i32, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _i32, _vp]),
    'escgnn_gemm_tf32x3': (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    'escgnn_gemm_tf32x3_bounded': (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _i32, _vp]),
    'escgnn_gemm_set_split_target': (_i32, [_i32]),
    'escgnn_linear_bn_fusable': (_i32, [_i32, _i32, _i32]),
    'escgnn_linear_bn_resident_ctas': (_i32, [_i32, _i32, _i32]),
    'escgnn_linear_bn_workspace_floats': (_i64, [_i32, _i32]),
    'escgnn_linear_bn_act_fwd': (_i32, [_vp, _i32, _vp, _i32, _vp, _i32, _i32, _i32] + [_vp] * 7 + [_i32, ctypes.c_float, ctypes.c_float,
                                        _vp, _i32, _vp, _i32, _vp, _i64, _vp]),
    'escgnn_linear_bn_act_bwd': (_i32, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp,
                                        _vp, _i32, _vp, _i64, _vp]),
    'escgnn_gemm_set_plan': (_i32, [_i32]),
    'escgnn_gemm_set_wide': (_i32, [_i32]),
    'escgnn_gemm_set_trace': (_i32, [_vp]),
    'escgnn_gemm_trace_slots': (_i32, []),
    'escgnn_bn_set_trace': (_i32, [_vp]),
    'escgnn_gemm_set_split_warps': (_i32, [_i32]),
    'escgnn_gemm_set_staged_store': (_i32, [_i32]),
    'escgnn_gemm_set_kb_groups': (_i32, [_i32]),
    'escgnn_gemm_set_drain': (_i32, [_i32]),
    'escgnn_gemm_workspace_floats': (_i64, [_i32, _i32, _i32]),
    'escgnn_tf32_split_lo': (_i32, [_vp, _i32, _vp, _i32, _i64, _i32, _vp]),
    'escgnn_gemm_simple': (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp]),
    'escgnn_adam_step_device': (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'escgnn_make_dims': (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    'escgnn_p2p_alloc': (_i32, [_i64, ctypes.POINTER(ctypes.c_void_p), _vp]),
    'escgnn_p2p_open': (_i32, [_vp, ctypes.POINTER(ctypes.c_void_p)]),
    'escgnn_p2p_close': (_i32, [_vp]),
    'escgnn_p2p_free': (_i32, [_vp]),
    'escgnn_p2p_flag_words': (_i64, [_i32]),
    'escgnn_allreduce_adam': (_i32, [_vp, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    'escgnn_allreduce_adam_range': (_i32, [_vp, _vp, _vp, _i32, _i32, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    'escgnn_adam_step': (_i32, [_vp, _vp, _vp, _vp, _i64] + [ctypes.c_float] * 4 + [_i64, ctypes.c_float, _vp]),
    'escgnn_all_pairs_spd_smem_bytes': (_i64, [_i64, _i64]),
    'escgnn_all_pairs_spd': (_i32, [_vp] * 4 + [_i64, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    'escgnn_edge_distance': (_i32, [_vp, _i32, _vp, _vp, _i64, _i32, _i32, ctypes.c_float, _vp, _vp, _vp, _vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('esc_gnn_b200: %s is missing -- build it with `python -m esc_gnn_b200.build`; '
                               'there is no CPU fallback' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError = header and library disagree: fail loudly
            fn.restype = res
            fn.argtypes = args
        if os.environ.get('ESCGNN_SPLIT_TARGET'):
            L.escgnn_gemm_set_split_target(int(os.environ['ESCGNN_SPLIT_TARGET']))
        if os.environ.get('ESCGNN_GEMM_DRAIN'):
            L.escgnn_gemm_set_drain(int(os.environ['ESCGNN_GEMM_DRAIN']))
        if os.environ.get('ESCGNN_GEMM_SPLIT_WARPS'):        # A/B switch: 4 or 8 splitter / epilogue warps
            L.escgnn_gemm_set_split_warps(int(os.environ['ESCGNN_GEMM_SPLIT_WARPS']))
        if os.environ.get('ESCGNN_GEMM_KB_GROUPS'):          # A/B switch: 1 or 2 k-block groups of splitter warps
            L.escgnn_gemm_set_kb_groups(int(os.environ['ESCGNN_GEMM_KB_GROUPS']))
        if os.environ.get('ESCGNN_GEMM_STAGED', '1') == '0':   # A/B switch: accumulator rows stored straight from registers
            L.escgnn_gemm_set_staged_store(0)
        if os.environ.get('ESCGNN_GEMM_WIDE', '1') == '0':   # A/B switch: 128-wide tiles everywhere
            L.escgnn_gemm_set_wide(0)
        if os.environ.get('ESCGNN_CLUSTER_BN', '1') == '0':
            L.escgnn_set_cluster_bn(0)
        if os.environ.get('ESCGNN_PDL', '1') == '0':      # A/B switch: plain stream-ordered launches
            L.escgnn_set_pdl(0)
        _lib = L
    return _lib


# kernels launched by one successful call of each entry point (bench.py's `gpu_launches` evidence)
KERNELS_PER_CALL = {'rewrite_self_loops': 5, 'encode_rd': 1, 'encode': 1, 'encode_subset': 1, 'scan': 3, 'expand_records': 1,
                    'csr_build': 4, 'sorted_ids_to_ptr': 1, 'bag_embed_fwd': 1, 'bag_embed_bwd': 1,
                    'gine_aggregate_fwd': 1, 'gine_aggregate_bwd': 2, 'gine_aggregate_fwd_ld': 1, 'gine_aggregate_bwd_ld': 2, 'gine_aggregate_bwd_ld_noeps': 1, 'segment_pool_fwd': 1, 'segment_pool_bwd': 1,
                    'edge_distance': 2, 'all_pairs_spd': 1, 'adam_step': 1, 'collate_edges': 1, 'ptr_to_ids': 1, 'bn_act_fwd': 1, 'bn_act_bwd': 1,
                    'act_fwd': 1, 'act_bwd': 1, 'colsum': 1, 'embedding_fwd': 1, 'embedding_bwd': 1, 'loss_fwd_bwd': 1,
                    'make_dims': 1, 'adam_step_device': 2, 'bag_embed_bwd_sorted': 4, 'bag_index_build': 3, 'bag_embed_bwd_indexed': 1, 'reduce_sum': 1, 'zero_tail_rows': 1, 'gemm_tf32x3': 1, 'tf32_split_lo': 1, 'gemm_simple': 1,
                    'linear_bn_act_fwd': 1, 'linear_bn_act_bwd': 1, 'allreduce_adam': 1}
LAUNCHES = {'n': 0}
PROFILE = None      # bench.py: a list; every mark() appends (label, cuda event) -> per-kernel durations by differencing
PROFILE_EXTERNAL = False   # record graph-capturable ("external") events: per-kernel device times of a REPLAYED graph


def mark(label):
    if PROFILE is not None:
        import torch
        ev = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        ev.record()
        PROFILE.append((label, ev))


def check(rc, what):
    if rc == 0:
        LAUNCHES['n'] += KERNELS_PER_CALL.get(what, 1)
        if PROFILE is not None:
            mark(what)
        return
    names = {-1: 'bad argument', -2: 'graph too large for the kernels', -3: 'record capacity', -4: 'data error'}
    if rc < 0:
        raise RuntimeError('esc_gnn_b200.%s failed: %s' % (what, names.get(rc, rc)))
    raise RuntimeError('esc_gnn_b200.%s failed: CUDA error %d' % (what, rc))


def raise_data_errors(bits):
    if not bits:
        return
    msgs = [m for b, m in ERR_BITS.items() if bits & b]
    raise RuntimeError('esc_gnn_b200 encoder: ' + '; '.join(msgs))
