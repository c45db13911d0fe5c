"""Collation with the reference's rules (/root/reference/batch.py:25-149, called from dataloader.py:26-29).

Four rules matter on the efficient path: every key whose name contains 'index' is shifted by the running node
count (`Data.__inc__`, :112-113) and concatenated along the last dim; `pos_batch` is shifted by the running
`pos_batch.max()+1` (:70-71); `pos_enc` / `pos_index` are concatenated unshifted (:72-73); a `batch` vector maps
nodes to graphs (:120-123).  Offsets are applied with one vectorised add per key instead of one per graph.
"""
import torch

from .data import Data


def _increment(d, key, item):
    """What the reference adds to `cumsum[key]` after graph `d` (batch.py:68-113)."""
    if key == 'pos_batch':
        return int(item.max()) + 1
    if key in ('pos_enc', 'pos_index', 'edge_pos'):
        return 0
    if key == 'node_to_subgraph':
        return d.num_subgraphs
    if key in ('subgraph_to_graph', 'batch_2', 'batch_3'):
        return 1
    if key in ('original_edge_index', 'original_idx'):
        return d.original_num_nodes if hasattr(d, 'original_num_nodes') else 0
    if key == 'batch_edge':
        return d.original_edge_index.size()[1] if hasattr(d, 'original_edge_index') else 0
    if key in ('tree_edge_index', 'atom2clique_index', 'edge_index_2', 'edge_index_3', 'assignment2_to_subgraph',
               'assignment3_to_subgraph', 'assignment_index_2', 'assignment_index_3'):
        raise NotImplementedError('collation of %r belongs to the k-GNN / junction-tree baselines (SURVEY.md: out of scope)' % key)
    return d.__inc__(key, item)


class Batch(Data):
    def __init__(self, batch=None, **kwargs):
        super(Batch, self).__init__(**kwargs)
        self.batch = batch
        self.__data_class__ = Data
        self.__slices__ = None

    @staticmethod
    def from_data_list(data_list, follow_batch=[]):
        """Same result as the reference loop (batch.py:25-149): per key, every graph's tensor is shifted by the running
        increment of the graphs BEFORE it that carry the key, then concatenated along `__cat_dim__`; `<key>_batch` vectors for
        the keys in `follow_batch` (:43-44,114-116); `batch` from the graphs whose `num_nodes` is known (:118-123).  The
        shifts are applied with one vectorised add per key instead of one tensor add per graph and key."""
        keys = set()
        for d in data_list:
            keys |= set(d.keys)
        keys = sorted(keys)
        assert 'batch' not in keys
        out = Batch()
        out.__data_class__ = data_list[0].__class__
        slices = {k: [0] for k in keys}
        cols = {k: [] for k in keys}                 # per key: the items of the graphs that carry it, in order
        incs = {k: [] for k in keys}                 # per key: each of those graphs' increment (reference `cumsum[key] += ...`)
        owner = {k: [] for k in keys}                # per key: index of the graph each item came from
        node_counts = []
        for i, d in enumerate(data_list):
            node_counts.append(d.num_nodes)
            for k in d.keys:
                item = d[k]
                if torch.is_tensor(item) and item.dim() == 0:      # (the reference raises on 0-d tensors; accepted here as [1])
                    item = item.view(1)
                cols[k].append(item)
                owner[k].append(i)
                size = item.size(d.__cat_dim__(k, item)) if torch.is_tensor(item) else 1
                slices[k].append(slices[k][-1] + size)
                incs[k].append(_increment(d, k, item))
        for k in keys:
            items = cols[k]
            first = items[0]
            if torch.is_tensor(first):
                dim = data_list[0].__cat_dim__(k, first)
                cat = torch.cat(items, dim=dim)
                inc = incs[k]
                if first.dtype != torch.bool and any(torch.is_tensor(v) or v != 0 for v in inc[:-1]):
                    if any(torch.is_tensor(v) and v.dim() > 0 for v in inc):
                        raise NotImplementedError('collation of %r (per-row increments) is outside the efficient path' % k)
                    sizes = torch.tensor([t.size(dim) for t in items], dtype=torch.long)
                    off = torch.tensor([0] + [int(v) for v in inc[:-1]], dtype=torch.long).cumsum(0)
                    cat = cat + torch.repeat_interleave(off, sizes).to(device=cat.device, dtype=cat.dtype)
                out[k] = cat
            elif isinstance(first, (int, float)):
                out[k] = torch.tensor(items)
            else:
                out[k] = items
            if k in follow_batch:
                sizes = torch.tensor([slices[k][j + 1] - slices[k][j] for j in range(len(items))], dtype=torch.long)
                out['%s_batch' % k] = torch.repeat_interleave(torch.tensor(owner[k], dtype=torch.long), sizes)
        if node_counts[-1] is None:                  # reference :125-126 looks at the LAST graph only
            out.batch = None
        else:
            known = [(i, int(n)) for i, n in enumerate(node_counts) if n is not None]
            out.batch = torch.repeat_interleave(torch.tensor([i for i, _ in known], dtype=torch.long),
                                                torch.tensor([n for _, n in known], dtype=torch.long))
        out.__slices__ = slices
        out.__num_graphs__ = len(data_list)
        return out.contiguous()

    @property
    def num_graphs(self):
        n = self.__dict__.get('__num_graphs__')
        if n is not None:
            return n
        return int(self.batch[-1]) + 1          # reference: batch.py:214-217
