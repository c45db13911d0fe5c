"""CPU, gloo, world_size 2: sharding of the encoder's units and the data-parallel gradient exchange."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from esc_gnn_b200 import distributed as D
from esc_gnn_b200 import synth


def test_shard_bounds_cover_and_balance():
    _, _, eptr, nptr = synth.make_batch_arrays(5, 0, 64)
    cost = D.encoder_cost(eptr, nptr, 3)
    for world in (1, 2, 4, 8):
        b = D.shard_bounds(cost, world)
        assert b[0] == 0 and b[-1] == 64 and (np.diff(b) >= 0).all()
        loads = [cost[b[i]:b[i + 1]].sum() for i in range(world)]
        assert max(loads) <= 1.0 * cost.sum() / world + cost.max()
    assert D.shard_bounds([], 4).tolist() == [0, 0, 0, 0, 0]


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    _, _, eptr, nptr = synth.make_batch_arrays(2, 0, 40)
    lo, hi = D.my_shard(eptr, nptr, 3)
    owned = torch.zeros(40, dtype=torch.int64)
    owned[lo:hi] = 1
    dist.all_reduce(owned)                                   # every graph owned by exactly one rank, no exchange of data
    g = torch.Generator().manual_seed(rank)
    grad = torch.randn(1000, generator=g)
    mine = grad.clone()
    D.allreduce_mean_(grad)
    both = [torch.randn(1000, generator=torch.Generator().manual_seed(r)) for r in range(world)]
    ok = bool((owned == 1).all()) and torch.allclose(grad, sum(both) / world, atol=1e-6) and not torch.equal(grad, mine)
    if rank == 0:
        out.put(ok)
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_gradient_mean():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok and all(p.exitcode == 0 for p in procs)
