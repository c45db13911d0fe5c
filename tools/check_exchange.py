"""Data-parallel exchange check, run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_exchange.py
The same model / batches through the engine with exchange='nccl' and exchange='p2p' (csrc/p2p.cu): losses agree, the p2p replicas
are bit-identical across ranks, and the step times of both are printed (CUDA events, max over ranks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from esc_gnn_b200.pipeline import RawBatch

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
torch.backends.cuda.matmul.allow_tf32 = False

class A(object):
    pipeline = 1; encoder_ctas = 74; fuse_bn = 0; exchange = 'nccl'

pool = [RawBatch.synth(bench.CONFIG, rank * 1_000_000 + i * bench.BATCH, bench.BATCH).cuda(non_blocking=False) for i in range(6)]
out = {}
for mode in ('nccl', 'p2p-one', 'p2p'):
    A.exchange = mode.split('-')[0]
    eng = bench.build_engine(bench.CONFIG, 'zinc', bench.BATCH, pool, world, A)
    eng.bucketed_exchange = mode == 'p2p'
    losses = []
    for i in range(8):
        l = eng.step(pool[i % 6])
        if l is not None:
            losses.append(float(l.item()))
    eng.check_errors()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 200
    a.record()
    for i in range(K):
        eng.step(pool[i % 6])
    b.record(); torch.cuda.synchronize()
    eng.check_errors()
    t = torch.tensor([a.elapsed_time(b) / K], device='cuda', dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    flat = eng.opt.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = bool((flat == ref).all())
    ok = torch.tensor([int(same)], device='cuda'); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out[mode] = (losses, float(t[0]), bool(ok[0]), flat)
    if rank == 0:
        print('%-8s ms/step %.4f  graphs/s %.0f  replicas identical %s  losses %s' % (mode, float(t[0]), bench.BATCH * world / float(t[0]) * 1e3,
                                                                                   bool(ok[0]), ['%.5f' % v for v in losses[:6]]), flush=True)
if rank == 0:
    la, lb = out['nccl'][0], out['p2p'][0]
    d = max(abs(x - y) / max(1.0, abs(x)) for x, y in zip(la[:4], lb[:4]))
    d2 = max(abs(x - y) / max(1.0, abs(x)) for x, y in zip(out['p2p-one'][0][:6], lb[:6]))
    print('loss agreement nccl vs p2p (rel): %.2e, one bucket vs two: %.2e' % (d, d2), 'PASS' if d < 2e-3 and d2 < 2e-3 and out['p2p'][2] and out['p2p-one'][2] else 'FAIL')
dist.destroy_process_group()
