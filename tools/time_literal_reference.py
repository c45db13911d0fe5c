"""Time the LITERAL reference (unmodified source under /root/reference) on the bench workload, in the build container.

    python tools/time_literal_reference.py [--config 2] [--graphs 256] [--steps 5] [--procs P]

* extraction: `utils_edge_efficient.create_subgraphs` (utils_edge_efficient.py:20-152), imported as-is under the PyG
  stand-in of tests/_pyg_shim, over a `multiprocessing.Pool(P)` with one torch thread per worker (the
  north-star's "networkx/pqdm extraction" path in spirit -- a process pool over graphs, README.md:75);
* collation: the reference's own `batch.Batch.from_data_list` (batch.py:25-149);
* train step: the `NestedGIN_eff` class AST-extracted from zinc_models.py:504-611 / run_graphcount.py:39-194 /
  `GNN` from ogb_mol_gnn.py, forward + loss + backward + Adam on `torch.set_num_threads(P)`, median of `--steps` steps
  after one warm-up.

The reference cannot travel to the GPU box (no torch_geometric there either, and /root/reference is not mounted), so
this number is recorded once, labelled "build box", in BASELINE.md and in bench.py's `cpu_baseline.literal_reference`.
Prints one JSON line.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', '_pyg_shim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))


def _encode_one(job):
    import torch
    torch.set_num_threads(1)
    from torch_geometric.data import Data
    import utils_edge_efficient as REF
    g, fl = job
    kw = dict(x=torch.as_tensor(g['x']), edge_index=torch.as_tensor(g['edge_index'], dtype=torch.long), y=torch.as_tensor(g['y']).view(-1) if torch.as_tensor(g['y']).dim() == 0 else torch.as_tensor(g['y']))
    if 'edge_attr' in g:
        kw['edge_attr'] = torch.as_tensor(g['edge_attr'])
    d = Data(**kw)
    d.num_nodes = g['num_nodes']
    return REF.create_subgraphs(d, fl['h'], use_rd=fl['use_rd'], self_loop=fl['self_loop'])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', type=int, default=2)
    ap.add_argument('--graphs', type=int, default=256)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--procs', type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    import torch
    from esc_gnn_b200 import synth
    from tests import model_util as MU
    import make_golden_model as GM
    import batch as REF_BATCH
    fl = synth.ENCODER_FLAGS[args.config]
    graphs = [synth.make_graph(args.config, 10_000_000 + i) for i in range(args.graphs)]
    jobs = [(g, fl) for g in graphs]
    with mp.get_context('fork').Pool(args.procs) as pool:
        pool.map(_encode_one, jobs[:args.procs])                  # warm the workers (imports)
        t0 = time.perf_counter()
        datas = pool.map(_encode_one, jobs, chunksize=max(1, args.graphs // (4 * args.procs)))
        t_enc = time.perf_counter() - t0
    torch.set_num_threads(args.procs)
    t0 = time.perf_counter()
    batch = REF_BATCH.Batch.from_data_list(datas)
    t_col = time.perf_counter() - t0
    variant = {1: 'count', 2: 'zinc', 3: 'count', 4: 'ogb'}[args.config]
    kw = {'count': dict(num_layers=5, hidden=256), 'zinc': dict(num_layers=5),
          'ogb': dict(num_tasks=1, num_layer=6, emb_dim=300, drop_ratio=0.65, virtual_node=True, residual=False)}[variant]
    torch.manual_seed(0)
    model = GM.build_reference_model(variant, kw)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    b = batch                                    # the reference's own Batch object goes straight into the reference's class
    times = []
    for i in range(args.steps + 1):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = MU.loss_fn(variant, model(b), b.y)
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    t_trn = statistics.median(times[1:])
    n = args.graphs
    out = dict(what='unmodified reference source under the PyG stand-in (tests/_pyg_shim), build container',
               config=args.config, graphs=n, procs=args.procs, cpu=os.cpu_count(),
               encode_graphs_per_s=n / t_enc, encode_s=t_enc, collate_s=t_col, train_step_s=t_trn, train_graphs_per_s=n / t_trn,
               step_graphs_per_s=n / (t_enc + t_col + t_trn), nodes=int(batch.x.shape[0]), edges=int(batch.edge_index.shape[1]),
               nnz=int(batch.pos_enc.numel()), loss=float(loss.item()), torch=torch.__version__)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
