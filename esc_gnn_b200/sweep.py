"""Encoding-extraction sweep (BASELINE.json configs[4]): graphs of 25-500 nodes, h = 1..4, with and without appended
self-loops, processed in chunks of thousands of graphs per launch and sharded by chunk over the ranks with no
communication -- the batched replacement of the reference's per-graph `pre_transform` loop
(GraphCountDataset.py:111-117 -> utils_edge_efficient.py:20-152; run_ogb_mol.py:329-332).

Every chunk goes through `encode_batch(..., expand=True)`, i.e. the reference contract (rewritten edge_index + int64
pos_enc / pos_index / pos_batch) is materialised in HBM, and is reduced to (edges, records, sum of counts) so sweeps of any
length fit.  `digest()` is an order-independent checksum of an encoded batch, defined so that the CPU checker's batch driver
(outside this package) computes the same four numbers; the caller (bench.py / tests) compares a prefix of the sweep with it.
"""
import numpy as np
import torch

from . import synth
from .transform import encode_batch

_M1, _M2, _M3 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xC2B2AE3D27D4EB4F), np.uint64(0xBF58476D1CE4E5B9)


def digest(edge_ptr, num_edges, pos_enc, pos_index, pos_batch):
    """(E_out, nnz, sum of counts, xor-hash) of one encoded batch (host numpy; pos_batch = batch-wide edge ordinal)."""
    ep = np.asarray(edge_ptr, dtype=np.int64)
    pe, pi, pb = (np.asarray(a, dtype=np.int64) for a in (pos_enc, pos_index, pos_batch))
    g = np.searchsorted(ep, pb, side='right') - 1                       # graph of every record
    local = (pb - ep[g]).astype(np.uint64)
    with np.errstate(over='ignore'):
        x = (g.astype(np.uint64) + np.uint64(1)) * _M1 ^ local * _M2 ^ (pi.astype(np.uint64) << np.uint64(32)) ^ pe.astype(np.uint64)
        x ^= x >> np.uint64(29)
        x *= _M3
        x ^= x >> np.uint64(32)
    return int(num_edges), int(pe.shape[0]), int(pe.sum()), int(np.bitwise_xor.reduce(x).astype(np.int64)) if x.size else 0


def tiled_chunk(config, pool, chunk, start=0):
    """`chunk` graphs as packed arrays: `pool` distinct synthetic graphs of `config` (synth.make_graph(config, start + i)) tiled."""
    src, dst, eptr, nptr = synth.make_batch_arrays(config, start, pool)
    reps = (chunk + pool - 1) // pool
    if reps == 1 and pool == chunk:
        return src, dst, eptr, nptr
    de, dn = np.tile(np.diff(eptr), reps), np.tile(np.diff(nptr), reps)
    csrc, cdst = np.tile(src, reps), np.tile(dst, reps)
    ceptr = np.concatenate([[0], np.cumsum(de)])[:chunk + 1]
    cnptr = np.concatenate([[0], np.cumsum(dn)])[:chunk + 1]
    return csrc[:ceptr[-1]], cdst[:ceptr[-1]], ceptr.astype(np.int64), cnptr.astype(np.int64)


def run_sweep(total_graphs, chunk=8192, pool=2048, hs=(1, 2, 3, 4), loops=(False, True), world=1, rank=0, config=5, use_rd=False,
              check_graphs=0, reduce_max=None, reduce_sum=None):
    """Encode `total_graphs` graphs (ceil to whole chunks; chunk c belongs to rank c % world) for every (h, self_loop).

    Returns {'h{h}_loops{0|1}': dict(graphs_per_s, edges_per_s, records, contract_GBps, ms_per_chunk, digest_prefix)}; times are
    CUDA-event times of the rank's whole loop, `reduce_max` / `reduce_sum` (callables on python floats, e.g. NCCL all-reduces)
    fold them over ranks.  `digest_prefix` = digest() of the first `check_graphs` graphs (rank 0 only)."""
    reduce_max = reduce_max or (lambda v: v)
    reduce_sum = reduce_sum or (lambda v: v)
    src, dst, eptr, nptr = tiled_chunk(config, pool, chunk)
    n_chunks = (total_graphs + chunk - 1) // chunk
    mine = [c for c in range(n_chunks) if c % world == rank]
    ds, dd = torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda()
    te, tn = torch.as_tensor(eptr), torch.as_tensor(nptr)
    out = {}
    for h in hs:
        for sl in loops:
            pref = None
            if check_graphs and rank == 0:
                k = min(check_graphs, chunk)
                r = encode_batch(ds[:eptr[k]], dd[:eptr[k]], te[:k + 1], tn[:k + 1], h, use_rd, sl)
                pref = digest(r.edge_ptr.cpu().numpy(), r.num_edges, r.pos_enc.cpu().numpy(), r.pos_index.cpu().numpy(),
                              r.pos_batch.cpu().numpy())
                del r
            r = encode_batch(ds, dd, te, tn, h, use_rd, sl)          # warm-up (allocator, kernel attributes)
            del r
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            edges = recs = 0
            total = torch.zeros((), dtype=torch.int64, device='cuda')
            a.record()
            for _ in mine:
                r = encode_batch(ds, dd, te, tn, h, use_rd, sl)
                edges += r.num_edges; recs += r.nnz
                total += r.pos_enc.sum()
                del r
            b.record()
            torch.cuda.synchronize()
            ms = reduce_max(a.elapsed_time(b)) if mine else reduce_max(0.0)
            g_all, e_all, r_all = reduce_sum(float(len(mine) * chunk)), reduce_sum(float(edges)), reduce_sum(float(recs))
            e_in_all = reduce_sum(float(len(mine) * int(eptr[-1])))
            sec = max(ms, 1e-9) * 1e-3
            out['h%d_loops%d' % (h, int(sl))] = dict(
                graphs=g_all, graphs_per_s=g_all / sec, edges_per_s=e_all / sec, records=r_all, sum_counts=float(total),
                contract_GBps=(16 * e_in_all + 16 * e_all + 24 * r_all) / sec / 1e9, ms_per_chunk=ms / max(len(mine), 1),
                digest_prefix=pref)
    return out
