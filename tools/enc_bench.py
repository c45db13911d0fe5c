"""Developer benchmark of the encoder kernels (per-kernel CUDA-event times, several shapes). Not the judged bench."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esc_gnn_b200 import synth  # noqa: E402
from esc_gnn_b200.transform import encode_batch, encode_batch_host  # noqa: E402


def tiled(config, pool, graphs):
    src, dst, eptr, nptr = synth.make_batch_arrays(config, 0, pool)
    reps = (graphs + pool - 1) // pool
    return (np.tile(src, reps), np.tile(dst, reps), np.concatenate([[0], np.cumsum(np.tile(np.diff(eptr), reps))]),
            np.concatenate([[0], np.cumsum(np.tile(np.diff(nptr), reps))]))


def run(config, graphs, h=None, use_rd=None, self_loop=None, iters=5, pool=256):
    fl = dict(synth.ENCODER_FLAGS[config])
    if h is not None: fl['h'] = h
    if use_rd is not None: fl['use_rd'] = use_rd
    if self_loop is not None: fl['self_loop'] = self_loop
    src, dst, eptr, nptr = tiled(config, pool, graphs)
    G = len(nptr) - 1
    ds, dd = torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda()
    te, tn = torch.as_tensor(eptr), torch.as_tensor(nptr)
    mn = int(np.diff(nptr).max()); me = int((np.diff(eptr) + (np.diff(nptr) if fl['self_loop'] else 0)).max())
    for _ in range(2):
        r = encode_batch(ds, dd, te, tn, **fl)
    torch.cuda.synchronize()
    tim = {}
    t0 = time.perf_counter()
    for _ in range(iters):
        r = encode_batch(ds, dd, te, tn, timings=tim, **fl)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / iters
    ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in tim.items()}
    E_in, E_out, nnz = len(src), r.num_edges, r.nnz
    bytes_alg = 16 * E_in + 16 * E_out + 24 * nnz
    out = dict(config=config, graphs=G, flags=fl, E_out=E_out, nnz=nnz, ms=ms, wall_ms=wall * 1e3,
               graphs_per_s_kernels=G / (sum(ms.values()) * 1e-3), graphs_per_s_wall=G / wall,
               alg_GBps_kernels=bytes_alg / (sum(ms.values()) * 1e-3) / 1e9)
    t0 = time.perf_counter()
    rh = encode_batch_host(src, dst, eptr, nptr, **fl)
    out['host_e2e_graphs_per_s'] = G / (time.perf_counter() - t0)
    print(json.dumps(out))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--graphs', type=int, default=8192)
    a = ap.parse_args()
    for cfg in (2, 1, 4):
        run(cfg, a.graphs)
        run(cfg, a.graphs, use_rd=False)
    for h in (1, 2, 3, 4):
        run(5, 1024, h=h, pool=128)
