"""Batched `pre_transform`: what the reference's dataset `process()` loops do one graph at a time
(/root/reference/GraphCountDataset.py:111-117, dataset_zinc.py:76-85: `data_list = [pre_transform(d) for d in data_list]`),
done in chunks of thousands of graphs per kernel launch.  Returns per-graph `Data` objects with exactly the fields the
per-graph `create_subgraphs` produces, so the result can be collated / cached like the reference's processed dataset.
"""
import numpy as np
import torch

from .transform import encode_batch_host


def pre_transform_batched(data_list, h=1, use_rd=False, self_loop=False, chunk=4096, device=None):
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
        # pos_batch comes back as the BATCH-wide edge ordinal (ascending), so the records of graph g are exactly those with
        # edge_ptr[g] <= pos_batch < edge_ptr[g+1]; the graph-local ordinal the per-graph transform returns is pos_batch - edge_ptr[g]
        r = encode_batch_host(src, dst, eptr, nptr, h, use_rd, self_loop, local_ordinals=False, device=device)
        ep = r.edge_ptr.numpy()
        rec_ptr = np.searchsorted(r.pos_batch.numpy(), ep)      # records of graph g: [rec_ptr[g], rec_ptr[g+1])
        for g, d in enumerate(part):
            a, b, ra, rb = ep[g], ep[g + 1], rec_ptr[g], rec_ptr[g + 1]
            edge_attr = d.edge_attr
            if self_loop and edge_attr is not None:
                keep = d.edge_index[0] != d.edge_index[1]
                edge_attr = edge_attr[keep]
                edge_attr = torch.cat([edge_attr, edge_attr.new_full((nn[g], ) + tuple(edge_attr.size()[1:]), 1.)], dim=0)
            out.append(d.__class__(d.x, r.edge_index[:, a:b].clone(), edge_attr, d.y, None, pos_enc=r.pos_enc[ra:rb].clone(),
                                   pos_index=r.pos_index[ra:rb].clone(), pos_batch=r.pos_batch[ra:rb] - int(a)))
    return out


# ------------------------------------------------------------------------------------------------------------------
# Processed-dataset cache: the `(data, slices)` layout of `InMemoryDataset.collate` that the reference's `process()`
# methods write with `torch.save` (GraphCountDataset.py:118-119, dataset_zinc.py:87-88, dataset_pyg.py:183-186).  On disk the
# first element is a plain {key: tensor} dict rather than a pickled PyG `Data` (loadable with torch.load's weights-only
# default and without torch_geometric); `Data(**store)` rebuilds the reference's object, the slices dict is identical.
def collate(data_list):
    """All graphs concatenated key by key (along `Data.__cat_dim__`) + per-key boundary offsets.  No index increments
    (those belong to batching, batch.py), no `batch` vector."""
    from .data import Data
    keys = list(data_list[0].keys)
    big, slices = Data(), {}
    for key in keys:
        items = [d[key] for d in data_list]
        if torch.is_tensor(items[0]) and items[0].dim() > 0:
            dim = data_list[0].__cat_dim__(key, items[0])
            sizes = [it.size(dim) for it in items]
            big[key] = torch.cat(items, dim=dim)
        else:                                       # python numbers / 0-d tensors: one entry per graph
            sizes = [1] * len(items)
            big[key] = torch.as_tensor([it.item() if torch.is_tensor(it) else it for it in items])
        slices[key] = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64))
    nn = []
    for d in data_list:
        n = d.num_nodes
        nn.append(int(n.item()) if torch.is_tensor(n) else int(n))
    big['_num_nodes'] = torch.as_tensor(nn, dtype=torch.int64)          # kept explicitly: isolated tail nodes must survive
    slices['_num_nodes'] = torch.arange(len(nn) + 1, dtype=torch.int64)
    return big, slices


def separate(big, slices, idx, cls=None):
    """Graph `idx` of a collated store as its own `Data` (views into the store, like InMemoryDataset.get)."""
    from .data import Data
    out = (cls or Data)()
    for key in big.keys:
        a, b = int(slices[key][idx]), int(slices[key][idx + 1])
        v = big[key]
        if key == '_num_nodes':
            out.num_nodes = int(v[a])
            continue
        dim = out.__cat_dim__(key, v) if v.dim() > 1 else 0
        out[key] = v.narrow(dim if dim >= 0 else v.dim() + dim, a, b - a)
    return out


class EncodedDataset(object):
    """`InMemoryDataset`-style dataset whose `process()` runs the B200 encoder over the whole raw graph list in batched
    launches and caches the result under `root/processed/` (reference flow: GraphCountDataset.py:97-120,
    dataset_zinc.py:64-88 -- per-graph `pre_transform` loop, then `collate`, then `torch.save`).

    raw_graphs: a list of `Data` or a zero-argument callable returning one (read only when the cache is missing).
    The cache file name carries the encoder flags, like the reference's `path += '_h3_rd'` conventions
    (run_graphcount.py:395-401)."""

    def __init__(self, root, raw_graphs, h, use_rd=False, self_loop=False, transform=None, pre_filter=None, chunk=4096,
                 device=None, name='data'):
        import os
        self.root, self.transform = root, transform
        self.flags = dict(h=h, use_rd=bool(use_rd), self_loop=bool(self_loop))
        tag = '%s_h%s%s%s.pt' % (name, h, '_rd' if use_rd else '', '_loop' if self_loop else '')
        self.processed_path = os.path.join(root, 'processed', tag)
        if not os.path.exists(self.processed_path):
            graphs = raw_graphs() if callable(raw_graphs) else list(raw_graphs)
            if pre_filter is not None:
                graphs = [g for g in graphs if pre_filter(g)]
            graphs = pre_transform_batched(graphs, h=h, use_rd=use_rd, self_loop=self_loop, chunk=chunk, device=device)
            os.makedirs(os.path.dirname(self.processed_path), exist_ok=True)
            big, slices = collate(graphs)
            torch.save(({k: big[k] for k in big.keys}, slices), self.processed_path)
        store, self.slices = torch.load(self.processed_path)
        from .data import Data
        self.data = Data(**store)
        self._indices = None

    def __len__(self):
        return len(self._indices) if self._indices is not None else int(self.slices['_num_nodes'].numel()) - 1

    def get(self, idx):
        return separate(self.data, self.slices, idx)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            i = int(idx)
            if i < 0:
                i += len(self)
            d = self.get(i if self._indices is None else int(self._indices[i]))
            return self.transform(d) if self.transform is not None else d
        return self.index_select(idx)

    def index_select(self, idx):
        base = np.arange(len(self)) if self._indices is None else np.asarray(self._indices)
        if isinstance(idx, slice):
            sel = base[idx]
        else:
            idx = torch.as_tensor(idx)
            sel = base[idx.numpy()] if idx.dtype != torch.bool else base[idx.numpy().astype(bool)]
        out = self.__class__.__new__(self.__class__)
        out.__dict__.update(self.__dict__)
        out._indices = np.asarray(sel)
        return out

    def shuffle(self, generator=None):
        return self.index_select(torch.randperm(len(self), generator=generator))

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]
