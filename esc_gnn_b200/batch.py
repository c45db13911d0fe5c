"""Collation with the reference's rules (/root/reference/batch.py:25-149, called from dataloader.py:26-29).

Four rules matter on the efficient path: every key whose name contains 'index' is shifted by the running node
count (`Data.__inc__`, :112-113) and concatenated along the last dim; `pos_batch` is shifted by the running
`pos_batch.max()+1` (:70-71); `pos_enc` / `pos_index` are concatenated unshifted (:72-73); a `batch` vector maps
nodes to graphs (:120-123).  Offsets are applied with one vectorised add per key instead of one per graph.
"""
import torch

from .data import Data


class Batch(Data):
    def __init__(self, batch=None, **kwargs):
        super(Batch, self).__init__(**kwargs)
        self.batch = batch
        self.__data_class__ = Data
        self.__slices__ = None

    @staticmethod
    def from_data_list(data_list, follow_batch=[]):
        keys = set()
        for d in data_list:
            keys |= set(d.keys)
        keys = sorted(keys)
        assert 'batch' not in keys
        out = Batch()
        out.__data_class__ = data_list[0].__class__
        slices = {k: [0] for k in keys}
        cols = {k: [] for k in keys}
        node_counts, pb_counts = [], []
        for d in data_list:
            n = d.num_nodes
            node_counts.append(int(n) if n is not None else 0)
            for k in d.keys:
                item = d[k]
                cols[k].append(item)
                size = item.size(d.__cat_dim__(k, item)) if torch.is_tensor(item) and item.dim() > 0 else 1
                slices[k].append(slices[k][-1] + size)
            if 'pos_batch' in d:
                pb_counts.append(int(d['pos_batch'].max()) + 1)
        node_off = torch.tensor([0] + node_counts[:-1], dtype=torch.long).cumsum(0)
        for k in keys:
            items = cols[k]
            first = items[0]
            if torch.is_tensor(first):
                if first.dim() == 0:
                    items = [t.view(1) for t in items]
                    first = items[0]
                dim = data_list[0].__cat_dim__(k, first)
                cat = torch.cat(items, dim=dim)
                sizes = torch.tensor([t.size(dim) for t in items], dtype=torch.long)
                if first.dtype != torch.bool:
                    if k == 'pos_batch':
                        off = torch.tensor([0] + pb_counts[:-1], dtype=torch.long).cumsum(0)
                        cat = cat + torch.repeat_interleave(off, sizes).to(cat.device)
                    elif k in ('pos_enc', 'pos_index', 'edge_pos'):
                        pass
                    elif 'index' in k:
                        cat = cat + torch.repeat_interleave(node_off, sizes).to(device=cat.device, dtype=cat.dtype)
                out[k] = cat
            elif isinstance(first, (int, float)):
                out[k] = torch.tensor(items)
            else:
                out[k] = items
        if all(n is not None for n in node_counts):
            out.batch = torch.repeat_interleave(torch.arange(len(data_list)), torch.tensor(node_counts))
        out.__slices__ = slices
        out.__num_graphs__ = len(data_list)
        return out.contiguous()

    @property
    def num_graphs(self):
        n = self.__dict__.get('__num_graphs__')
        if n is not None:
            return n
        return int(self.batch[-1]) + 1          # reference: batch.py:214-217
