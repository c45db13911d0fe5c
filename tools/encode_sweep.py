"""BASELINE.json config 5: encoding-extraction sweep over synthetic graphs of 25-500 nodes, h = 1..4, rd off, sharded by
graph over the visible GPUs with no communication (one process per GPU under torchrun, or a single process).

    python tools/encode_sweep.py --graphs 65536 --chunk 8192 [--check 256]
Outputs are reduced per chunk to (edges, records, sum of counts) so arbitrarily long sweeps fit; the first --check graphs
are compared record-for-record with the C oracle.  The graph pool is generated once (synth.make_graph(5, i)) and tiled.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esc_gnn_b200 import distributed as D, synth  # noqa: E402
from esc_gnn_b200.transform import encode_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--graphs', type=int, default=65536)
    ap.add_argument('--pool', type=int, default=2048)
    ap.add_argument('--chunk', type=int, default=8192)
    ap.add_argument('--check', type=int, default=128)
    a = ap.parse_args()
    world, rank = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    if world > 1:
        torch.distributed.init_process_group('nccl')
    src, dst, eptr, nptr = synth.make_batch_arrays(5, 0, a.pool)
    reps = (a.chunk + a.pool - 1) // a.pool
    csrc, cdst = np.tile(src, reps), np.tile(dst, reps)
    ceptr = np.concatenate([[0], np.cumsum(np.tile(np.diff(eptr), reps))])
    cnptr = np.concatenate([[0], np.cumsum(np.tile(np.diff(nptr), reps))])
    n_chunks = (a.graphs + a.chunk - 1) // a.chunk
    mine = [c for c in range(n_chunks) if c % world == rank]          # chunks are independent units: round-robin shards
    ds, dd = torch.as_tensor(csrc).cuda(), torch.as_tensor(cdst).cuda()
    te, tn = torch.as_tensor(ceptr), torch.as_tensor(cnptr)
    out = {}
    for h in (1, 2, 3, 4):
        for sl in (False, True):
            if a.check and rank == 0:
                from oracle import c_oracle
                k = min(a.check, a.pool)
                r = encode_batch(ds[:eptr[k]], dd[:eptr[k]], torch.as_tensor(eptr[:k + 1]), torch.as_tensor(nptr[:k + 1]), h, False, sl)
                base = 0
                pe, pi, pb = r.pos_enc.cpu().numpy(), r.pos_index.cpu().numpy(), r.pos_batch.cpu().numpy()
                for g in range(k):
                    w = c_oracle.encode_graph(np.stack([src[eptr[g]:eptr[g + 1]], dst[eptr[g]:eptr[g + 1]]]), int(nptr[g + 1] - nptr[g]), h, False, sl)
                    m = (pb >= base) & (pb < base + w[0].shape[1])
                    assert np.array_equal(pe[m], w[1]) and np.array_equal(pi[m], w[2]), (h, sl, g)
                    base += w[0].shape[1]
            encode_batch(ds, dd, te, tn, h, False, sl)          # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            edges = recs = total = 0
            for _ in mine:
                r = encode_batch(ds, dd, te, tn, h, False, sl)
                edges += r.num_edges; recs += r.nnz; total += int(r.pos_enc.sum())
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], device='cuda', dtype=torch.float64)
            if world > 1:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            g_done = len(mine) * a.chunk
            gt = torch.tensor([g_done, edges, recs], device='cuda', dtype=torch.float64)
            if world > 1:
                torch.distributed.all_reduce(gt)
            if rank == 0:
                gps = float(gt[0]) / float(t[0])
                bytes_alg = 16 * float(gt[1]) + 16 * float(gt[1]) + 24 * float(gt[2])
                out['h%d_loops%d' % (h, int(sl))] = dict(graphs_per_s=gps, edges_per_s=float(gt[1]) / float(t[0]),
                                                          records=float(gt[2]), contract_GBps=bytes_alg / float(t[0]) / 1e9)
    if rank == 0:
        print(json.dumps(dict(workload='config 5 sweep: n~U{25..500}, m=1.25n, rd off', graphs=a.graphs, n_gpus=world,
                              oracle_checked_graphs=a.check, results=out)))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
