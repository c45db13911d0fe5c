"""One eager engine step of the bench workload (config 2, batch 256) between cudaProfilerStart/Stop, for
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_step python tools/profile_step.py
Every kernel of the step (encoder, collation, bag-embed, GINE aggregation, BatchNorm, tcgen05 GEMMs, pooling, readout, Adam) is
captured once, in launch order, on ONE stream (inline branches); ncu's cache control flushes the caches before every replay pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from esc_gnn_b200.pipeline import RawBatch

class A(object):
    pipeline = 0; encoder_ctas = 0; fuse_bn = int(os.environ.get('FUSE_BN', '0'))

torch.backends.cuda.matmul.allow_tf32 = False
pool = [RawBatch.synth(bench.CONFIG, i * bench.BATCH, bench.BATCH).cuda(non_blocking=False) for i in range(2)]
eng = bench.build_engine(bench.CONFIG, 'zinc', bench.BATCH, pool, 1, A, pipeline=False)
eng.use_graph = False
eng.inline_branches = True
for i in range(3):
    eng.step(pool[i % 2])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loss = eng.step(pool[1])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('loss', float(loss.item()), 'dims', eng.c.dims.tolist())
