"""Batched `pre_transform`: what the reference's dataset `process()` loops do one graph at a time
(/root/reference/GraphCountDataset.py:111-117, dataset_zinc.py:76-85: `data_list = [pre_transform(d) for d in data_list]`),
done in chunks of thousands of graphs per kernel launch.  Returns per-graph `Data` objects with exactly the fields the
per-graph `create_subgraphs` produces, so the result can be collated / cached like the reference's processed dataset.
"""
import numpy as np
import torch

from .transform import encode_batch_host


def pre_transform_batched(data_list, h=1, use_rd=False, self_loop=False, chunk=4096, device=0):
    """Equivalent to `[create_subgraphs(d, h, use_rd=use_rd, self_loop=self_loop) for d in data_list]`."""
    out = []
    for lo in range(0, len(data_list), chunk):
        part = data_list[lo:lo + chunk]
        nn = []
        for d in part:
            n = d.num_nodes
            nn.append(int(n.item()) if torch.is_tensor(n) else int(n))
        ee = [int(d.edge_index.size(1)) for d in part]
        src = torch.cat([d.edge_index[0] for d in part]).to(torch.int64)
        dst = torch.cat([d.edge_index[1] for d in part]).to(torch.int64)
        eptr = np.concatenate([[0], np.cumsum(ee)]).astype(np.int64)
        nptr = np.concatenate([[0], np.cumsum(nn)]).astype(np.int64)
        r = encode_batch_host(src, dst, eptr, nptr, h, use_rd, self_loop, local_ordinals=True, device=device)
        ep = r.edge_ptr.numpy()
        rec_ptr = np.searchsorted(_global_ordinal(r.pos_batch.numpy(), ep), ep)      # records of graph g: [rec_ptr[g], rec_ptr[g+1])
        for g, d in enumerate(part):
            a, b, ra, rb = ep[g], ep[g + 1], rec_ptr[g], rec_ptr[g + 1]
            edge_attr = d.edge_attr
            if self_loop and edge_attr is not None:
                keep = d.edge_index[0] != d.edge_index[1]
                edge_attr = edge_attr[keep]
                edge_attr = torch.cat([edge_attr, edge_attr.new_full((nn[g], ) + tuple(edge_attr.size()[1:]), 1.)], dim=0)
            out.append(d.__class__(d.x, r.edge_index[:, a:b].clone(), edge_attr, d.y, None, pos_enc=r.pos_enc[ra:rb].clone(),
                                   pos_index=r.pos_index[ra:rb].clone(), pos_batch=r.pos_batch[ra:rb].clone()))
    return out


def _global_ordinal(pos_batch_local, edge_ptr):
    """pos_batch comes back graph-local (local_ordinals); rebuild the batch-wide edge ordinal to split records by graph.
    Records are grouped by graph and ascending inside a graph, so a graph boundary is where the local ordinal drops."""
    if pos_batch_local.size == 0:
        return pos_batch_local
    drops = np.nonzero(np.diff(pos_batch_local) < 0)[0] + 1
    # a graph whose first edge ordinal is not smaller than the previous graph's last one cannot happen: every graph
    # starts at local ordinal 0 and every edge has at least three records
    starts = np.concatenate([[0], drops])
    graph_of = np.zeros(pos_batch_local.size, dtype=np.int64)
    graph_of[starts[1:]] = 1
    graph_of = np.cumsum(graph_of)
    nonempty = np.nonzero(np.diff(edge_ptr) > 0)[0]
    return pos_batch_local + edge_ptr[nonempty[graph_of]]
