"""Headline benchmark: graphs/sec for ego-net encoding + NestedGIN_eff train step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic input: a batch of 256 ZINC-shaped raw graphs
(BASELINE.json configs[1]; run_zinc.py:56 batch size) is structurally encoded (h=3, rd on, no self-loops;
run_zinc.py:141-146), collated and run through one NestedGIN_eff train step (5 layers, hidden 256; forward, L1
loss, backward, Adam).  `value` = graphs/s with the raw graphs already resident in HBM; `e2e` = the same step fed
from pinned HOST buffers (H2D of the raw graphs and D2H of the loss inside the timed region).
Multi-GPU: weak scaling, every rank encodes and trains on its own 256 graphs; the only exchange is the gradient exchange of the
data-parallel step: reduce-scatter + Adam + all-gather fused into one kernel over NVLink peer memory (`config.exchange` = 'p2p';
'nccl' = one NCCL all-reduce of the flat gradient between two captured graphs, the fallback).

Beside the headline the same JSON line carries (none of them is the judged `value`):
  configs          the other BASELINE.json model configs through the same engine: cfg 1 (count_cycle-shaped, batch 128, h=3),
                   cfg 3 (count_graphlet-shaped, batch 32, h=4), cfg 4 (ogbg-molhiv-shaped GNN gin_eff, batch 32, h=4)
  extraction       encode_batch on 8192 ZINC-shaped graphs per rank, device-resident I/O, reference contract materialised
  extraction_e2e   the same through the host front end: host arrays in -> host arrays out (pipelined, compact and int64 forms)
  sweep            BASELINE.json configs[4]: >= 1 M graphs of 25-500 nodes, h = 1..4, both loop modes, sharded by chunk over the
                   ranks (strong scaling: the total is fixed), a 1000-graph prefix checked against the C oracle's digest
  dropin           the literal drop-in call sequence: per-graph create_subgraphs -> DataLoader -> model(batch) train step
  roofline / cpu_baseline / sequential / large_batch   as documented in DESIGN.md section 5

`--impl reference` times the CPU restatement of the same step (oracle port: C encoder with OpenMP over graphs +
plain-PyTorch NestedGIN_eff on all host threads); the unmodified Python reference cannot travel to the GPU box (its
build-box timing is in BASELINE.md and quoted as `cpu_baseline.literal_reference`).
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
# the other BASELINE.json model configs (config id -> engine variant, batch, description)
OTHER = {1: ('count', 128, 'count_cycle-shaped (n~19, 31 edges), h=3 rd loops, NestedGIN_eff 5 layers hidden 256, batch 128 (run_graphcount.py:465)'),
         3: ('count', 32, 'count_graphlet-shaped, h=4 rd loops, NestedGIN_eff 5 layers hidden 256, batch 32'),
         4: ('ogb', 32, 'ogbg-molhiv-shaped (n~26), h=4 rd loops, GNN gin_eff 6 layers emb 300 virtual node dropout 0.65, batch 32 (run_ogb_mol.py:432-434)')}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops_sustained', 1400.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1400.0, 'fallback (B200_PROFILING.md)'


def load_json(rel):
    path = os.path.join(ROOT, rel)
    return json.load(open(path)) if os.path.exists(path) else None


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

    def mark(self):
        return len(self.rows)

    def summary(self, lo=0, hi=None):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        rows = self.rows[lo:hi]
        sm = [float(r[1]) for r in rows if len(r) >= 8 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith('active')})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_step_factory(threads):
    """The oracle port of one step on the host: C encoder (OpenMP over graphs) + torch CPU train step."""
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


def literal_reference():
    """Build-box timing of the UNMODIFIED reference on the same 256-graph batches (tools/time_literal_reference.py)."""
    d = load_json('profiles/r02_literal_reference_buildbox.json')
    if not d:
        return None
    row = [r for r in d['rows'] if r['config'] == CONFIG]
    if not row:
        return None
    r = row[0]
    return dict(where='build box (no GPU), %d cores: static number, not measured in this run' % r['procs'], graphs=r['graphs'],
                encode_graphs_per_s=r['encode_graphs_per_s'], train_graphs_per_s=r['train_graphs_per_s'],
                step_graphs_per_s=r['step_graphs_per_s'], source='profiles/r02_literal_reference_buildbox.json',
                what='unmodified utils_edge_efficient.create_subgraphs (process pool) + batch.py collation + AST-extracted '
                     'zinc_models.NestedGIN_eff train step under the PyG stand-in')


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
                                 encode_graphs_per_s=sample * args.steps / t_enc, train_graphs_per_s=sample * args.steps / t_trn,
                                 literal_reference=literal_reference()),
               e2e=dict(value=value, unit='graphs/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------- own arm (B200)
def build_model(variant):
    import torch
    from esc_gnn_b200 import graphcount_model, ogb_model, zinc_model
    torch.manual_seed(0)
    if variant == 'zinc':
        m = zinc_model.NestedGIN_eff(None, LAYERS)
    elif variant == 'count':
        m = graphcount_model.NestedGIN_eff(None, LAYERS, HIDDEN, use_rd=True, graph_pred=False, dropout=0, edge_nest=True, use_cycle=True)
    else:
        m = ogb_model.GNN('ogbg-molhiv', 1, num_layer=6, emb_dim=300, gnn_type='gin_eff', virtual_node=True, residual=False,
                          drop_ratio=0.65)
    return m.cuda().train()


def build_engine(config, variant, batch, host_pool, world, args, pipeline=None):
    from esc_gnn_b200 import synth
    from esc_gnn_b200.engine import StaticTrainEngine
    fl = synth.ENCODER_FLAGS[config]
    nodes_cap = int(max(b.num_nodes for b in host_pool) * 1.04) + 64
    edges_cap = int(max(b.src.numel() for b in host_pool) * 1.04) + 128
    mx_n = max(40, max(b.max_nodes for b in host_pool))
    mx_e = max(96, max((b.max_loop_edges if fl['self_loop'] else b.max_in_edges) for b in host_pool))
    pipe = bool(args.pipeline) if pipeline is None else pipeline
    return StaticTrainEngine(build_model(variant), variant, fl, max_graphs=batch, max_nodes_per_graph=mx_n, max_edges_per_graph=mx_e,
                             nodes_cap=nodes_cap, edges_cap=edges_cap, lr=LR, distributed=world > 1, use_graph=True, pipeline=pipe,
                             encoder_ctas=args.encoder_ctas if pipe else None, fuse_bn=bool(args.fuse_bn),
                             exchange=getattr(args, 'exchange', 'auto'))


def zinc_flops(n_nodes, e_out, graphs):
    """Useful (fp32-equivalent) flops of the Linear layers of one FORWARD pass; dgrad and wgrad are the same again."""
    L1, H = LAYERS - 1, HIDDEN
    return 2.0 * (e_out * H * H + e_out * (H + 32) * (32 + L1 * H) + n_nodes * (32 * H + H * H) + L1 * n_nodes * 2 * H * H +
                  graphs * (LAYERS * H * H + H))


def zinc_gemm_bytes(n_nodes, e_out, graphs):
    """Algorithmic operand bytes (A + B read, C written, fp32) of the Linear layers of one forward pass; dgrad and wgrad move the
    same three matrices in other roles, so each role's bytes equal this."""
    L1, H = LAYERS - 1, HIDDEN
    lin = [(e_out, H, H), (e_out, H + 32, 32 + L1 * H), (n_nodes, 32, H), (n_nodes, H, H)] + [(n_nodes, H, H)] * (2 * L1) + \
          [(graphs, LAYERS * H, H), (graphs, H, 1)]
    return 4.0 * sum(m * k + n * k + m * n for m, k, n in lin)


def run_own(args):
    import torch
    import torch.distributed as dist
    from esc_gnn_b200 import _lib, synth
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
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')     # > 126 MB L2
    W = max(args.warmup, 3)

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

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(t)
        return float(t[0])

    def engine_numbers(eng, dev_pool, host_pool, k_steps, batch):
        """(graphs/s value, graphs/s e2e, ms/step value, ms/step e2e) of one engine, whole job."""
        n_pool = len(host_pool)
        for i in range(W + 2):                              # W >= 3 warm-up steps (+2 eager steps before the graph capture)
            eng.step(dev_pool[i % n_pool])
            read(eng.step(host_pool[i % n_pool]))
        eng.check_errors()
        barrier()
        ms_v = timed(lambda b: eng.step(b), dev_pool, k_steps)
        barrier()
        ms_v = max_over_ranks(ms_v)
        barrier()
        # e2e: inputs from pinned host memory every step, and every step's loss read by the host -- through step_read(), which copies
        # the 4 bytes to a pinned slot behind the step and hands the host the PREVIOUS step's value (its event has fired), so that
        # reading the result does not idle the device once per step; the last loss is read inside the timed function's final sync
        ms_e = timed(lambda b: eng.step_read(b), host_pool, k_steps)
        last = eng.last_read()
        assert last is None or last == last, 'loss is NaN'
        barrier()
        ms_e = max_over_ranks(ms_e)
        eng.check_errors()
        g = batch * world * k_steps
        return g / (ms_v * 1e-3), g / (ms_e * 1e-3), ms_v / k_steps, ms_e / k_steps

    clocks = ClockSampler(local)        # sampled from the warm-up to the end of the timed regions (the GPU is under load throughout)
    clocks.start()
    clocks.wait_first()
    # ================================================================== headline: config 2, batch 256
    n_pool = min(args.steps + args.warmup, 12)
    host_pool = [RawBatch.synth(CONFIG, rank * 1_000_000 + i * BATCH, BATCH) for i in range(n_pool)]
    dev_pool = [b.cuda(non_blocking=False) for b in host_pool]
    eng = build_engine(CONFIG, 'zinc', BATCH, host_pool, world, args)
    c0 = clocks.mark()
    value, e2e_value, ms_step, ms_step_e2e = engine_numbers(eng, dev_pool, host_pool, args.steps, BATCH)
    c1 = clocks.mark()
    # ---- the same engine without the encoder/training overlap, for comparison (reported, not the headline)
    sequential = None
    if args.pipeline:
        seq = build_engine(CONFIG, 'zinc', BATCH, host_pool, world, args, pipeline=False)
        for i in range(W + 2):
            seq.step(dev_pool[i % n_pool])
        barrier()
        ms_seq = timed(lambda b: seq.step(b), dev_pool, args.steps)
        barrier()
        ms_seq = max_over_ranks(ms_seq)
        seq.check_errors()
        sequential = dict(value=BATCH * world * args.steps / (ms_seq * 1e-3), unit='graphs/s', ms_per_step=ms_seq / args.steps)
        del seq
    # ================================================================== the other BASELINE model configs (1, 3, 4)
    configs = {}
    if not args.no_configs:
        k_c = min(args.steps, 100)
        for cfg, (variant, batch, what) in OTHER.items():
            hp = [RawBatch.synth(cfg, 7_000_000 + rank * 100_000 + i * batch, batch) for i in range(6)]
            dp = [b.cuda(non_blocking=False) for b in hp]
            e_c = build_engine(cfg, variant, batch, hp, world, args)
            v, e, ms_v, ms_e = engine_numbers(e_c, dp, hp, k_c, batch)
            d = e_c.c.dims.cpu().tolist()
            configs['cfg%d' % cfg] = dict(workload=what, value=v, unit='graphs/s', ms_per_step=ms_v, steps=k_c, batch_per_gpu=batch,
                                          e2e=dict(value=e, unit='graphs/s', ms_per_step=ms_e, h2d_bytes_per_step=hp[0].h2d_bytes(),
                                                   d2h_bytes_per_step=4),
                                          shape=dict(nodes=d[0], edges=d[1], nnz=d[3]))
            del e_c, dp, hp
    # ================================================================== extraction (hot path 1 alone)
    large, extraction, extraction_e2e, sweep_out, dropin = None, None, None, None, None
    LG = 8192
    if not args.no_large:
        # ---- device-resident I/O: every rank encodes its own 8192 graphs through the reference contract (int64 pos_enc / pos_index /
        # pos_batch + rewritten edge_index); whole-job graphs/s.  The encoder shards by graph with no collective.
        from esc_gnn_b200.transform import HostEncoder, encode_batch
        raw_lh = RawBatch.synth(CONFIG, 5_000_000 + rank * LG, LG)
        raw_l = raw_lh.cuda(non_blocking=False)
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
        nnz_l, e_l = r_enc.nnz, r_enc.num_edges
        del r_enc
        # ---- host arrays in -> host arrays out through the plugin call (C-ABI escgnn_encode_host_submit / _wait): pinned staging,
        # 4-byte records over PCIe, D2H of chunk k under the kernels of chunk k+1; wall clock (host work is part of it), max over ranks
        h_arrays = [a.numpy() for a in (raw_lh.src, raw_lh.dst, raw_lh.edge_ptr, raw_lh.node_ptr)]
        henc = HostEncoder(fl['h'], fl['use_rd'], fl['self_loop'])
        pinned = dict(pos_enc=torch.empty(nnz_l + 1024, dtype=torch.int64).pin_memory(),
                      pos_index=torch.empty(nnz_l + 1024, dtype=torch.int64).pin_memory(),
                      pos_batch=torch.empty(nnz_l + 1024, dtype=torch.int64).pin_memory())
        k_h = 12

        host_threads = max(1, (os.cpu_count() or 1) // world)      # the ranks of one box share its cores

        def host_run(expand):
            tot = 0
            for r in henc.stream(h_arrays for _ in range(k_h)):
                if expand:
                    r.expand(out=pinned, threads=host_threads)
                tot += r.nnz
            return tot
        res = {}
        for name, expand in (('compact', False), ('int64_triple', True)):
            host_run(expand)                                 # warm-up: arenas reach their final size
            barrier()
            t0 = time.perf_counter()
            tot = host_run(expand)
            sec = max_over_ranks(time.perf_counter() - t0)
            assert tot == k_h * nnz_l
            in_b = sum(a.nbytes for a in h_arrays)
            res[name] = dict(value=LG * world * k_h / sec, unit='graphs/s', ms_per_chunk=1e3 * sec / k_h, h2d_bytes_per_chunk=in_b,
                             d2h_bytes_per_chunk=4 * nnz_l + 12 * e_l, host_bytes_expanded_per_chunk=(24 * nnz_l if expand else 0))
        extraction_e2e = dict(chunk_graphs=LG, chunks=k_h, compact=res['compact'], int64_triple=res['int64_triple'],
                              what='HostEncoder.stream: numpy int64 arrays in -> pinned host arrays out; compact = records (index | count<<11) + '
                                   'per-edge offsets / counts (what crosses PCIe); int64_triple = plus the host-side expansion to the '
                                   'reference contract (escgnn_expand_records_host, %d host threads per rank)' % host_threads,
                              timing='wall clock incl. host staging copies, max over ranks')
        del henc, pinned
    # ---- config-5 sweep (strong scaling: the total number of graphs is fixed, chunks are dealt round-robin to the ranks)
    if not args.no_sweep:
        from esc_gnn_b200 import sweep
        check = args.sweep_check if rank == 0 else 0
        t0 = time.perf_counter()
        res = sweep.run_sweep(args.sweep_graphs, chunk=8192, pool=2048, world=world, rank=rank, check_graphs=check,
                              reduce_max=max_over_ranks, reduce_sum=sum_over_ranks)
        oracle_ok = None
        if rank == 0 and check:
            from oracle import c_oracle                      # the checker, never the thing measured
            src, dst, eptr, nptr = sweep.tiled_chunk(5, 2048, 8192)
            oracle_ok = True
            for key, v in res.items():
                h, sl = int(key[1]), key.endswith('1')
                want = c_oracle.encode_batch_digest(src[:eptr[check]], dst[:eptr[check]], eptr[:check + 1], nptr[:check + 1], h, False, sl)
                v['oracle_prefix_match'] = tuple(v['digest_prefix']) == tuple(want)
                oracle_ok = oracle_ok and v['oracle_prefix_match']
            if not oracle_ok:
                raise RuntimeError('config-5 sweep: GPU digest of the %d-graph prefix differs from the C oracle: %r' % (check, res))
        for v in res.values():
            v.pop('digest_prefix', None)
        sweep_out = dict(workload='BASELINE.json configs[4]: n ~ U{25..500}, m = floor(1.25 n), rd off; %d graphs = %d chunks of 8192 '
                                  '(2048 distinct graphs tiled), h = 1..4, without / with appended self-loops' % (
                                      args.sweep_graphs, (args.sweep_graphs + 8191) // 8192),
                         scaling='strong', n_gpus=world, oracle_checked_graphs=check if rank == 0 else None, oracle_prefix_match=oracle_ok,
                         results=res, wall_s=time.perf_counter() - t0,
                         what='encode_batch per chunk with the int64 triple materialised in HBM; graphs/s = all graphs / max-over-ranks CUDA-event time')
    # ---- the literal drop-in call sequence (what a user of the reference's scripts runs): per-graph create_subgraphs on host Data
    # objects -> DataLoader / Batch.from_data_list -> model(batch) with torch autograd around the kernels -> Adam
    if world == 1 and not args.no_dropin:
        from esc_gnn_b200 import DataLoader, create_subgraphs
        from esc_gnn_b200.data import Data
        from esc_gnn_b200.optim import FlatAdam
        graphs = [synth.make_graph(CONFIG, 20_000_000 + i) for i in range(2 * BATCH)]
        datas = [Data(x=torch.as_tensor(g['x']), edge_index=torch.as_tensor(g['edge_index']), edge_attr=torch.as_tensor(g['edge_attr']),
                      y=torch.as_tensor(g['y']).view(1)) for g in graphs]
        model_d = build_model('zinc')
        opt_d = FlatAdam(model_d.parameters(), lr=LR)
        t_enc = t_trn = lv = 0.0
        for rep in range(2):                                # first pass = warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            enc_d = [create_subgraphs(d, fl['h'], use_rd=fl['use_rd'], self_loop=fl['self_loop']) for d in datas[rep * BATCH:(rep + 1) * BATCH]]
            t1 = time.perf_counter()
            for batch in DataLoader(enc_d, batch_size=BATCH, shuffle=False):
                batch = batch.to('cuda')
                opt_d.zero_grad()
                loss = torch.nn.L1Loss()(model_d(batch), batch.y.view(-1, 1))
                loss.backward()
                opt_d.step(1)
                lv = float(loss.item())
            t2 = time.perf_counter()
            t_enc, t_trn = t1 - t0, t2 - t1
        dropin = dict(value=BATCH / (t_enc + t_trn), unit='graphs/s', graphs=BATCH, encode_graphs_per_s=BATCH / t_enc,
                      collate_train_graphs_per_s=BATCH / t_trn, loss=lv, timing='wall clock, second pass',
                      what='256 x create_subgraphs(Data) one graph per call (host tensors, C-ABI host front end) -> DataLoader -> '
                           'NestedGIN_eff(batch) + L1 + backward (torch autograd over the sm_100a kernels) + Adam')
        del model_d, opt_d
    # ---- large batch ("report both reference-batch and large-batch (8 192 graphs) numbers"): the same engine
    # at 32x the reference batch, where the kernels stop being launch-latency bound; explains the roofline, not the headline
    if world == 1 and not args.no_large:
        from esc_gnn_b200.engine import StaticTrainEngine
        eng_l = StaticTrainEngine(build_model('zinc'), 'zinc', fl, max_graphs=LG, max_nodes_per_graph=40, max_edges_per_graph=96,
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
        fl_l = 3 * zinc_flops(d_l[0], d_l[1], LG)
        g_ms_l = sum(km_l.get(k, 0.0) for k in ('gemm_fwd', 'gemm_dgrad', 'gemm_wgrad', 'linear_bn_act_fwd', 'linear_bn_act_bwd'))
        enc_ms_l = km_l.get('encode', 0.0) + km_l.get('encode_rd', 0.0)
        enc_bytes_l = 16 * d_l[1] + 16 * d_l[1] + 24 * d_l[3]
        large = dict(graphs=LG, value=LG * k_l / (ms_l * 1e-3), unit='graphs/s', ms_per_step=ms_l / k_l,
                     shape=dict(nodes=d_l[0], edges=d_l[1], nnz=d_l[3]),
                     gemm_useful_tflops=fl_l / (g_ms_l * 1e-3) / 1e12 if g_ms_l else None, gemm_ms_per_step=g_ms_l,
                     gemm_frac_of_tensor_peak=(fl_l / (g_ms_l * 1e-3) / 1e12 / peaks()[1]) if g_ms_l else None,
                     extraction_graphs_per_s=LG / (enc_ms_l * 1e-3) if enc_ms_l else None, extraction_ms_per_step=enc_ms_l,
                     extraction_contract_GBps=enc_bytes_l / (enc_ms_l * 1e-3) / 1e9 if enc_ms_l else None,
                     kernel_ms_per_step={k: round(v, 4) for k, v in sorted(km_l.items(), key=lambda kv: -kv[1])[:12]})
        del eng_l
    clocks.stop()
    clk = clocks.summary(c0, c1)                           # the headline's timed regions
    clk['whole_run'] = clocks.summary()
    # ---- per-kernel device times: the step captured once more on one stream with an event after every launch, replayed
    launches_a = _lib.LAUNCHES['n']
    kernel_ms, calls = eng.profile(dev_pool[0], reps=min(args.steps, 10), flush=flush)
    launches_per_step = _lib.LAUNCHES['n'] - launches_a
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ================================================================== roofline of the dominant hand-written kernel
    hbm_peak, tensor_peak, peak_src = peaks()
    dims = eng.c.dims.cpu().tolist()
    n_nodes, e_out, nnz = dims[0], dims[1], dims[3]
    e_in = e_out
    H2 = HIDDEN
    traffic_file = load_json('profiles/r02_traffic.json') or {}
    own = {k: v for k, v in kernel_ms.items() if not k.startswith(('memset', 'copy', 'misc', 'end', 'start'))}
    alg = {'encode_rd': 16 * e_in + 24 * e_out, 'encode': 16 * e_in + 16 * e_out + 24 * nnz,
           'bag_embed_fwd': 12 * nnz + 4 * e_out * H2, 'bag_embed_bwd_indexed': 12 * nnz + 4 * e_out * H2,
           'gine_aggregate_fwd_ld': e_out * (2 * 4 * H2 + 8) + 2 * 4 * n_nodes * H2,
           'gine_aggregate_bwd_ld_noeps': e_out * (3 * 4 * H2 + 8) + 4 * e_out * H2 + 2 * 4 * n_nodes * H2,
           'segment_pool_fwd': 4 * n_nodes * LAYERS * H2 + 4 * BATCH * LAYERS * H2,
           'segment_pool_bwd': 4 * n_nodes * LAYERS * H2 + 4 * BATCH * LAYERS * H2,
           'bn_act_fwd': 2 * 4 * n_nodes * H2, 'bn_act_bwd': 4 * 4 * n_nodes * H2}
    sum_ms = sum(kernel_ms.values())
    gemm_labels = [k for k in ('gemm_fwd', 'gemm_dgrad', 'gemm_wgrad', 'linear_bn_act_fwd', 'linear_bn_act_bwd') if k in kernel_ms]
    g_ms = sum(kernel_ms[k] for k in gemm_labels)
    top = max(own, key=own.get)
    fwd_flops = zinc_flops(n_nodes, e_out, BATCH)
    if g_ms >= own[top]:
        # dense contractions: one kernel template, three roles (forward, dgrad, wgrad); each role does the SAME useful flops
        # (SURVEY 8(d) forward formula), so frac = (roles x forward flops) / time of all roles / peak
        role_of = {'gemm_fwd': 'fwd', 'linear_bn_act_fwd': 'fwd', 'gemm_dgrad': 'dgrad', 'linear_bn_act_bwd': 'dgrad', 'gemm_wgrad': 'wgrad'}
        roles = {}
        for k in gemm_labels:
            r = roles.setdefault(role_of[k], dict(useful_flops=fwd_flops, ms=0.0, launches=0.0))
            r['ms'] += kernel_ms[k]; r['launches'] += calls[k]
        for r in roles.values():
            r['tflops'] = r['useful_flops'] / (r['ms'] * 1e-3) / 1e12
            r['frac'] = r['tflops'] / tensor_peak
        g_calls = sum(r['launches'] for r in roles.values())
        flops = fwd_flops * len(roles)
        achieved = flops / (g_ms * 1e-3) / 1e12
        tr = traffic_file.get('gemm')
        roofline = dict(bound='tensor', kernel='gemm_tf32x3_ts_kernel (%s)' % ' + '.join(gemm_labels), achieved=achieved, peak=tensor_peak,
                        unit='TFLOP/s', frac=achieved / tensor_peak,
                        traffic=(tr['dram_bytes_per_launch'] if tr else None), traffic_source=(tr['source'] if tr else None),
                        algorithmic_operand_bytes_per_launch=zinc_gemm_bytes(n_nodes, e_out, BATCH) * len(roles) / max(g_calls, 1),
                        peak_source=peak_src + ' dense bf16, sustained',
                        share_of_step=g_ms / sum_ms, algorithmic_flops_per_step=flops, launches_per_step=g_calls,
                        launch_ms=g_ms / max(g_calls, 1), frac_of_3xtf32_ceiling=achieved / (tensor_peak / 6.0), roles=roles,
                        note='useful fp32-equivalent flops: forward formula of SURVEY 8(d) x the roles present; the kernel issues 3 tf32 MMA '
                             'passes per product and tf32 runs at half the bf16 rate, so 1/6 of this peak is the ceiling of a 3xTF32 scheme')
    else:
        per_launch_ms = kernel_ms[top] / max(calls[top], 1)
        bytes_launch = alg.get(top, alg['encode'])
        achieved = bytes_launch / (per_launch_ms * 1e-3) / 1e9
        tr = traffic_file.get(top)
        roofline = dict(bound='hbm', kernel=top, achieved=achieved, peak=hbm_peak, unit='GB/s', frac=achieved / hbm_peak,
                        traffic=(tr['dram_bytes_per_launch'] if tr else None), traffic_source=(tr['source'] if tr else None),
                        peak_source=peak_src, share_of_step=kernel_ms[top] / sum_ms,
                        algorithmic_bytes_per_launch=bytes_launch, launch_ms=per_launch_ms)
    roofline['how'] = ('kernel times: graph replay of the same launch sequence on one stream with a CUDA event after every '
                       'launch (the timed step overlaps weight-gradient work on a second graph branch)')
    # HBM-bound kernels of the step, each against the measured copy bandwidth (algorithmic bytes of SURVEY 8(d) per launch of the
    # widest instance; `traffic` = dram bytes per launch from the committed ncu capture when there is one)
    hbm_rows = {}
    for k, b in alg.items():
        if k in kernel_ms and calls.get(k):
            ms_k = kernel_ms[k] / calls[k]
            tr = traffic_file.get(k)
            hbm_rows[k] = dict(algorithmic_bytes_per_launch=b, launch_ms=ms_k, achieved_GBps=b / (ms_k * 1e-3) / 1e9,
                               frac=b / (ms_k * 1e-3) / 1e9 / hbm_peak, launches_per_step=calls[k],
                               traffic=(tr['dram_bytes_per_launch'] if tr else None))
    roofline['hbm_kernels'] = hbm_rows
    enc_ms = kernel_ms.get('encode', 0) + kernel_ms.get('encode_rd', 0)
    roofline['encoder'] = dict(kernels='ego_rd + ego_encode', ms_per_step=enc_ms, contract_bytes=alg['encode'],
                               achieved_GBps=alg['encode'] / (enc_ms * 1e-3) / 1e9 if enc_ms else None,
                               frac_of_hbm=alg['encode'] / (enc_ms * 1e-3) / 1e9 / hbm_peak if enc_ms else None,
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
                   encode_graphs_per_s=sample * reps / t_enc, train_graphs_per_s=sample * reps / t_trn,
                   literal_reference=literal_reference())
    out = dict(metric=METRIC, value=value, unit='graphs/s', n_gpus=world, steps=args.steps,
               warmup=W, ms_per_step=ms_step, higher_is_better=True, scaling='weak',
               vs_baseline=None, dtype='int64+f64 (encode), f32 (model)', data='synthetic',
               config=dict(workload=WORKLOAD, global_batch=BATCH * world, parallelism='dp%d' % world,
                           l2='flushed between timed iterations (256 MB write)', lr=LR,
                           pipeline=('encoder of batch k overlaps training of batch k-1 (one encode + one train step per step); '
                                     'encoder grids capped at %d CTAs' % args.encoder_ctas if args.pipeline else 'off'),
                           fused_linear_bn=bool(args.fuse_bn), exchange=eng.exchange),
               clocks=clk,
               e2e=dict(value=e2e_value, unit='graphs/s', h2d_bytes_per_step=host_pool[0].h2d_bytes(), d2h_bytes_per_step=4,
                        ms_per_step=ms_step_e2e,
                        how='engine.step_read(RawBatch in pinned host memory): H2D of the raw batch and D2H of the 4-byte loss inside every '
                            'timed step; the host reads step k-1\'s loss (pinned slot + event) after launching step k'),
               gpu_launches=launches_per_step * args.steps, launches_per_step=launches_per_step,
               kernel_ms_per_step={k: round(v, 5) for k, v in sorted(kernel_ms.items(), key=lambda kv: -kv[1])},
               roofline=roofline, cpu_baseline=cpu, sequential=sequential, large_batch=large, configs=configs, extraction=extraction,
               extraction_e2e=extraction_e2e, sweep=sweep_out, dropin=dropin,
               shape=dict(graphs=BATCH, nodes=n_nodes, edges=e_out, nnz=nnz),
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
    ap.add_argument('--exchange', default='auto', choices=['auto', 'nccl', 'p2p'],
                    help='data-parallel exchange: the fused NVLink peer-memory kernel (p2p), one NCCL all-reduce between two graphs (nccl), or p2p with nccl as the fallback (auto)')
    ap.add_argument('--cpu-sample', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-large', action='store_true', help='skip the 8192-graph sections (extraction, extraction_e2e, large_batch)')
    ap.add_argument('--no-configs', action='store_true', help='skip BASELINE configs 1, 3, 4')
    ap.add_argument('--no-sweep', action='store_true', help='skip the config-5 extraction sweep')
    ap.add_argument('--no-dropin', action='store_true', help='skip the per-graph drop-in sequence')
    ap.add_argument('--sweep-graphs', type=int, default=1 << 20, help='graphs of the config-5 sweep (whole job)')
    ap.add_argument('--sweep-check', type=int, default=1000, help='prefix of the sweep compared with the C oracle (rank 0)')
    ap.add_argument('--encoder-ctas', type=int, default=74, help='pipelined engine: cap of the encoder grids (0 = fill the machine)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_own(args)


if __name__ == '__main__':
    main()
