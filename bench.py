"""Headline benchmark: graphs/sec for ego-net encoding + NestedGIN_eff train step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic input: a batch of 256 ZINC-shaped raw graphs
(BASELINE.json configs[1]; run_zinc.py:56 batch size) is structurally encoded (h=3, rd on, no self-loops;
run_zinc.py:141-146), collated and run through one NestedGIN_eff train step (5 layers, hidden 256; forward, L1
loss, backward, Adam).  `value` = graphs/s with the raw graphs already resident in HBM; `e2e` = the same step fed
from pinned HOST buffers (H2D of the raw graphs and D2H of the loss inside the timed region).
Multi-GPU: weak scaling, every rank encodes and trains on its own 256 graphs; the only exchange is one NCCL
all-reduce of the flat gradient per step.

`--impl reference` times the CPU restatement of the same step (oracle port: C encoder with OpenMP over graphs +
plain-PyTorch NestedGIN_eff on all host threads); the unmodified Python reference cannot travel to the GPU box.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIG, LAYERS, HIDDEN, BATCH, LR = 2, 5, 256, 256, 1e-3
METRIC = 'graphs/sec, ego-net encoding + NestedGIN_eff train step'
WORKLOAD = 'ZINC-shaped synthetic molecules (n~23, ~25 bonds), encode h=3 rd=on + NestedGIN_eff 5 layers hidden 256 train step, batch 256'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return json.load(open(path)).get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith('active')})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_step_factory(threads):
    """The oracle port of one step on the host: C encoder (OpenMP over graphs) + torch CPU train step."""
    import numpy as np
    import torch
    from esc_gnn_b200 import synth
    from oracle import c_oracle, model_ref
    from tests import model_util as MU
    torch.set_num_threads(threads)
    fl = synth.ENCODER_FLAGS[CONFIG]
    model = model_ref.NestedGINEffZinc(LAYERS, HIDDEN)
    opt = torch.optim.Adam(model.parameters(), lr=LR)
    model.train()
    cache = {}

    def step(i, sample):
        if i not in cache:                         # inputs prepared outside the timed region
            arr = synth.make_batch_arrays(CONFIG, 10_000_000 + i * sample, sample)
            cache[i] = (arr, MU.ref_batch(CONFIG, 10_000_000 + i * sample, sample))
        (src, dst, eptr, nptr), batch = cache[i]
        t0 = time.perf_counter()
        c_oracle.encode_batch_digest(src, dst, eptr, nptr, fl['h'], fl['use_rd'], fl['self_loop'], threads=threads)
        t1 = time.perf_counter()
        opt.zero_grad()
        loss = torch.nn.L1Loss()(model(batch), batch.y.view(-1, 1))
        loss.backward()
        opt.step()
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, loss.item()
    return step


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = BATCH if args.cpu_sample is None else args.cpu_sample
    step = cpu_step_factory(threads)
    steps_ref = min(args.steps, 40)                      # ~0.25 s per CPU step: keeps the default run within a minute
    for i in range(min(max(args.warmup, 1), 3)):
        step(i % 4, sample)
    t_enc = t_trn = 0.0
    args.steps = steps_ref
    for i in range(args.steps):
        a, b, _ = step(i % 4, sample)
        t_enc += a; t_trn += b
    total = t_enc + t_trn
    value = sample * args.steps / total
    out = dict(impl='reference', metric=METRIC, value=value, unit='graphs/s', n_gpus=args.gpus, steps=args.steps,
               warmup=args.warmup, ms_per_step=1e3 * total / args.steps, higher_is_better=True, scaling='weak',
               vs_baseline=None, dtype='int64+f64 (encode), f32 (model)', data='synthetic',
               config=dict(workload=WORKLOAD, note='CPU oracle port; each step is a bounded sample of %d graphs of the workload' % sample),
               cpu_baseline=dict(value=value, unit='graphs/s', cores=threads, kind='port',
                                 sample='%d steps x %d ZINC-shaped graphs: C-oracle encode (OpenMP) + torch CPU train step' % (args.steps, sample),
                                 encode_graphs_per_s=sample * args.steps / t_enc, train_graphs_per_s=sample * args.steps / t_trn),
               e2e=dict(value=value, unit='graphs/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------- own arm (B200)
def run_own(args):
    import torch
    import torch.distributed as dist
    from esc_gnn_b200 import _lib, ops, synth, zinc_model
    from esc_gnn_b200.engine import StaticTrainEngine
    from esc_gnn_b200.pipeline import RawBatch
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference)')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.backends.cuda.matmul.allow_tf32 = False          # fp32 parity with the reference's fp32 CPU path
    torch.backends.cudnn.allow_tf32 = False
    fl = synth.ENCODER_FLAGS[CONFIG]
    torch.manual_seed(0)
    model = zinc_model.NestedGIN_eff(None, LAYERS).cuda()
    model.train()
    n_pool = min(args.steps + args.warmup, 12)
    host_pool = [RawBatch.synth(CONFIG, rank * 1_000_000 + i * BATCH, BATCH) for i in range(n_pool)]
    dev_pool = [b.cuda(non_blocking=False) for b in host_pool]
    nodes_cap = int(max(b.num_nodes for b in host_pool) * 1.04) + 64
    edges_cap = int(max(b.src.numel() for b in host_pool) * 1.04) + 128
    eng = StaticTrainEngine(model, 'zinc', fl, max_graphs=BATCH, max_nodes_per_graph=40, max_edges_per_graph=96,
                            nodes_cap=nodes_cap, edges_cap=edges_cap, lr=LR, distributed=world > 1, use_graph=True,
                            pipeline=bool(args.pipeline), encoder_ctas=args.encoder_ctas, fuse_bn=bool(args.fuse_bn))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')     # > 126 MB L2

    def read(loss):                                         # pipelined engines return the previous batch's loss (None at first)
        return float(loss.item()) if loss is not None else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, pool, k_steps):
        evs = []
        for i in range(k_steps):
            flush.zero_()                                   # L2 flush between timed iterations (untimed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn(pool[i % len(pool)])
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    clocks = ClockSampler(local)        # sampled from the warm-up to the end of the timed regions (the GPU is under load throughout)
    clocks.start()
    clocks.wait_first()
    for i in range(max(args.warmup, 3) + 2):                # W >= 3 warm-up steps (+2 eager steps before the graph capture)
        eng.step(dev_pool[i % n_pool])
        read(eng.step(host_pool[i % n_pool]))
    eng.check_errors()
    # ---- value: inputs resident in HBM
    launches0 = _lib.LAUNCHES['n']
    barrier()
    ms_value = timed(lambda b: eng.step(b), dev_pool, args.steps)
    barrier()
    ms_value = max_over_ranks(ms_value)
    # ---- e2e: raw graphs in pinned host memory, loss read back every step
    barrier()
    ms_e2e = timed(lambda b: read(eng.step(b)), host_pool, args.steps)
    barrier()
    ms_e2e = max_over_ranks(ms_e2e)
    eng.check_errors()
    # ---- the same engine without the encoder/training overlap, for comparison (reported, not the headline)
    sequential = None
    if args.pipeline:
        torch.manual_seed(0)
        model_s = zinc_model.NestedGIN_eff(None, LAYERS).cuda()
        model_s.train()
        seq = StaticTrainEngine(model_s, 'zinc', fl, max_graphs=BATCH, max_nodes_per_graph=40, max_edges_per_graph=96,
                                nodes_cap=nodes_cap, edges_cap=edges_cap, lr=LR, distributed=world > 1, use_graph=True)
        for i in range(max(args.warmup, 3) + 2):
            seq.step(dev_pool[i % n_pool])
        barrier()
        ms_seq = timed(lambda b: seq.step(b), dev_pool, args.steps)
        barrier()
        ms_seq = max_over_ranks(ms_seq)
        seq.check_errors()
        sequential = dict(value=BATCH * world * args.steps / (ms_seq * 1e-3), unit='graphs/s', ms_per_step=ms_seq / args.steps)
        del seq, model_s
    # ---- large batch ("report both reference-batch and large-batch (8 192 graphs) numbers"): the same engine
    # at 32x the reference batch, where the kernels stop being launch-latency bound; explains the roofline, not the headline
    large, extraction = None, None
    LG = 8192
    if not args.no_large:
        # ---- extraction alone (the encoder shards by graph with no collective): every rank encodes its own 8192
        # graphs through the reference contract (int64 pos_enc / pos_index / pos_batch + rewritten edge_index); whole-job graphs/s
        from esc_gnn_b200.transform import encode_batch
        raw_l = RawBatch.synth(CONFIG, 5_000_000 + rank * LG, LG).cuda(non_blocking=False)
        ep_h, np_h = raw_l.edge_ptr_host, raw_l.node_ptr_host
        enc = lambda: encode_batch(raw_l.src, raw_l.dst, ep_h, np_h, fl['h'], fl['use_rd'], fl['self_loop'], expand=True)
        for _ in range(3):
            r_enc = enc()
        barrier()
        k_x = 10
        ms_x = timed(lambda _b: enc(), [None], k_x)
        barrier()
        ms_x = max_over_ranks(ms_x)
        b_enc = 16 * raw_l.src.numel() + 16 * r_enc.num_edges + 24 * r_enc.nnz
        extraction = dict(value=LG * world * k_x / (ms_x * 1e-3), unit='graphs/s', graphs_per_rank=LG, ms_per_call=ms_x / k_x,
                          contract_bytes_per_call=b_enc, contract_GBps_per_gpu=b_enc / (ms_x / k_x * 1e-3) / 1e9,
                          what='encode_batch (h=3, rd on): E1 + E5 + E2-E4 kernels + expansion to the int64 triple, one device->host '
                               'read of two counters per call')
        del r_enc
    if world == 1 and not args.no_large:
        torch.manual_seed(0)
        model_l = zinc_model.NestedGIN_eff(None, LAYERS).cuda()
        model_l.train()
        eng_l = StaticTrainEngine(model_l, 'zinc', fl, max_graphs=LG, max_nodes_per_graph=40, max_edges_per_graph=96,
                                  nodes_cap=raw_l.num_nodes + 64, edges_cap=raw_l.src.numel() + 128, lr=LR, use_graph=True,
                                  pipeline=bool(args.pipeline))
        for _ in range(6):
            eng_l.step(raw_l)
        torch.cuda.synchronize()
        k_l = 20
        ms_l = timed(lambda b: eng_l.step(b), [raw_l], k_l)
        eng_l.check_errors()
        km_l, calls_l = eng_l.profile(raw_l, reps=3, flush=flush)
        d_l = eng_l.c.dims.cpu().tolist()
        L1 = LAYERS - 1
        fl_l = 3 * 2.0 * (d_l[1] * HIDDEN * HIDDEN + d_l[1] * (HIDDEN + 32) * (32 + L1 * HIDDEN) + d_l[0] * (32 * HIDDEN + HIDDEN * HIDDEN) +
                          L1 * d_l[0] * 2 * HIDDEN * HIDDEN + LG * (LAYERS * HIDDEN * HIDDEN + HIDDEN))
        g_ms_l = sum(km_l.get(k, 0.0) for k in ('gemm_fwd', 'gemm_dgrad', 'gemm_wgrad'))
        enc_ms_l = km_l.get('encode', 0.0) + km_l.get('encode_rd', 0.0)
        enc_bytes_l = 16 * d_l[1] + 16 * d_l[1] + 24 * d_l[3]
        large = dict(graphs=LG, value=LG * k_l / (ms_l * 1e-3), unit='graphs/s', ms_per_step=ms_l / k_l,
                     shape=dict(nodes=d_l[0], edges=d_l[1], nnz=d_l[3]),
                     gemm_useful_tflops=fl_l / (g_ms_l * 1e-3) / 1e12 if g_ms_l else None, gemm_ms_per_step=g_ms_l,
                     extraction_graphs_per_s=LG / (enc_ms_l * 1e-3) if enc_ms_l else None, extraction_ms_per_step=enc_ms_l,
                     extraction_contract_GBps=enc_bytes_l / (enc_ms_l * 1e-3) / 1e9 if enc_ms_l else None,
                     kernel_ms_per_step={k: round(v, 4) for k, v in sorted(km_l.items(), key=lambda kv: -kv[1])[:12]})
        del eng_l, model_l, raw_l
    clk = clocks.stop()
    # ---- per-kernel device times: the step captured once more on one stream with an event after every launch, replayed
    launches_a = _lib.LAUNCHES['n']
    kernel_ms, calls = eng.profile(dev_pool[0], reps=min(args.steps, 10), flush=flush)
    launches_per_step = _lib.LAUNCHES['n'] - launches_a
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant hand-written kernel (B_enc = 16 E_in + 16 E_out + 24 nnz)
    peak, peak_src = peaks()
    dims = eng.c.dims.cpu().tolist()
    n_nodes, e_out, nnz = dims[0], dims[1], dims[3]
    e_in = e_out
    own = {k: v for k, v in kernel_ms.items() if not k.startswith(('memset', 'copy', 'misc'))}
    top = max(own, key=own.get)
    H2 = HIDDEN
    alg = {'encode_rd': 16 * e_in + 24 * e_out, 'encode': 16 * e_in + 16 * e_out + 24 * nnz,
           'bag_embed_fwd': 12 * nnz + 4 * e_out * H2, 'bag_embed_bwd': 12 * nnz + 4 * e_out * H2,
           'gine_aggregate_fwd': e_out * (2 * 4 * H2 + 8) + 2 * 4 * n_nodes * H2,
           'gine_aggregate_bwd': e_out * (3 * 4 * H2 + 8) + 4 * e_out * H2 + 2 * 4 * n_nodes * H2,
           'bn_act_fwd': 3 * 4 * e_out * H2, 'bn_act_bwd': 5 * 4 * e_out * H2}
    per_launch_ms = kernel_ms[top] / max(calls[top], 1)
    sum_ms = sum(kernel_ms.values())
    if top.startswith('gemm'):
        # dense contraction: useful flops of the Linear layers of one step, 1/3 each for fwd, dgrad, wgrad
        L1 = LAYERS - 1
        flops = 2.0 * (e_out * H2 * H2 + e_out * (H2 + 32) * (32 + L1 * H2) + n_nodes * (32 * H2 + H2 * H2) +
                       L1 * n_nodes * 2 * H2 * H2 + BATCH * (LAYERS * H2 * H2 + H2))
        pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('bf16_tflops_sustained', 1388.2) \
            if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 1400.0
        g_labels = [k for k in ('gemm_fwd', 'gemm_dgrad', 'gemm_wgrad') if k in kernel_ms]      # one kernel, three roles
        g_ms, g_calls = sum(kernel_ms[k] for k in g_labels), sum(calls[k] for k in g_labels)
        achieved = flops / (g_ms * 1e-3) / 1e12
        roofline = dict(bound='tensor', kernel='gemm_tf32x3_kernel (%s)' % ' + '.join(g_labels), achieved=achieved, peak=pk,
                        unit='TFLOP/s', frac=achieved / pk, traffic=None, peak_source=peak_src + ' dense bf16, sustained',
                        share_of_step=g_ms / sum_ms, algorithmic_flops_per_step=flops, launches_per_step=g_calls,
                        launch_ms=g_ms / max(g_calls, 1), frac_of_3xtf32_ceiling=achieved / (pk / 6.0),
                        traffic_ncu=dict(launch='gemm_tf32x3_ts_kernel<128,0,0,2,2>, forward 12800x288 -> 256', dram_read_bytes=15102464,
                                         dram_write_bytes=24064, algorithmic_operand_bytes=4 * (12800 * 288 + 256 * 288),
                                         source='profiles/r01_prof_dense_r1d_metrics.txt (ncu --set full); output tile stays in L2'),
                        note='useful fp32-equivalent flops; the kernel issues 3 tf32 MMA passes per product and tf32 runs at half '
                             'the bf16 rate, so 1/6 of this peak is the ceiling of a 3xTF32 scheme')
    else:
        bytes_launch = alg.get(top, alg['encode'])
        achieved = bytes_launch / (per_launch_ms * 1e-3) / 1e9
        roofline = dict(bound='hbm', kernel=top, achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                        traffic=None, peak_source=peak_src, share_of_step=kernel_ms[top] / sum_ms,
                        algorithmic_bytes_per_launch=bytes_launch, launch_ms=per_launch_ms)
    roofline['how'] = ('kernel times: graph replay of the same launch sequence on one stream with a CUDA event after every '
                       'launch (the timed step overlaps weight-gradient work on a second graph branch)')
    enc_ms = kernel_ms.get('encode', 0) + kernel_ms.get('encode_rd', 0)
    roofline['encoder'] = dict(kernels='ego_rd + ego_encode', ms_per_step=enc_ms, contract_bytes=alg['encode'],
                               achieved_GBps=alg['encode'] / (enc_ms * 1e-3) / 1e9 if enc_ms else None,
                               frac_of_hbm=alg['encode'] / (enc_ms * 1e-3) / 1e9 / peak if enc_ms else None,
                               note='issue-bound integer / fp64 kernels (profiles/): the byte roofline is not what limits them')
    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        step = cpu_step_factory(threads)
        sample = BATCH
        step(0, sample)
        t_enc = t_trn = 0.0
        reps = 3
        for i in range(reps):
            a, b, _ = step(i % 2, sample)
            t_enc += a; t_trn += b
        cpu = dict(value=sample * reps / (t_enc + t_trn), unit='graphs/s', cores=threads, kind='port',
                   sample='%d steps x %d ZINC-shaped graphs: C-oracle encode (OpenMP) + torch CPU NestedGIN_eff train step' % (reps, sample),
                   encode_graphs_per_s=sample * reps / t_enc, train_graphs_per_s=sample * reps / t_trn)
    launches = launches_per_step * args.steps
    graphs = BATCH * world * args.steps
    h2d = host_pool[0].h2d_bytes()
    out = dict(metric=METRIC, value=graphs / (ms_value * 1e-3), unit='graphs/s', n_gpus=world, steps=args.steps,
               warmup=max(args.warmup, 3), ms_per_step=ms_value / args.steps, higher_is_better=True, scaling='weak',
               vs_baseline=None, dtype='int64+f64 (encode), f32 (model)', data='synthetic',
               config=dict(workload=WORKLOAD, global_batch=BATCH * world, parallelism='dp%d' % world,
                           l2='flushed between timed iterations (256 MB write)', lr=LR,
                           pipeline=('encoder of batch k overlaps training of batch k-1 (one encode + one train step per step); '
                                     'encoder grids capped at %d CTAs' % args.encoder_ctas if args.pipeline else 'off')),
               clocks=clk,
               e2e=dict(value=graphs / (ms_e2e * 1e-3), unit='graphs/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                        ms_per_step=ms_e2e / args.steps),
               gpu_launches=launches, kernel_ms_per_step={k: round(v, 5) for k, v in sorted(kernel_ms.items(), key=lambda kv: -kv[1])},
               roofline=roofline, cpu_baseline=cpu, sequential=sequential, large_batch=large, extraction=extraction,
               shape=dict(graphs=BATCH, nodes=n_nodes, edges=e_out, nnz=nnz, nodes_cap=nodes_cap, edges_cap=edges_cap),
               engine='one CUDA graph per step (encode+collate+fwd+bwd+Adam), programmatic dependent launches; every GEMM on the hand-written '
                      'tcgen05 3xTF32 kernel')
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--pipeline', type=int, default=1, help='1: overlap the encoder of batch k with the training of batch k-1')
    ap.add_argument('--fuse-bn', type=int, default=0, help='1: Linear+BatchNorm+act as one launch (GEMM epilogue behind a grid barrier)')
    ap.add_argument('--cpu-sample', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-large', action='store_true', help='skip the 8192-graph section')
    ap.add_argument('--encoder-ctas', type=int, default=74, help='pipelined engine: cap of the encoder grids (0 = fill the machine)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_own(args)


if __name__ == '__main__':
    main()
