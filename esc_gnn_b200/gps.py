"""GraphGPS twin of the hot path: what the reference adds to GraphGPS to use the structural encodings.

* `ESCEdgeEncoding` -- the `z_initial` / `z_embedding` members of `GPSLayer` and the first statement of its forward
  (GraphGPS/graphgps/layer/gps_layer.py:169-188): `batch.edge_attr += z_embedding(bag_embed(pos_enc, pos_index, pos_batch))`, on the
  sm_100a bag-embed / BatchNorm / tcgen05 Linear kernels.  Parameter names match the reference members (`z_initial.weight`,
  `z_embedding.{1,3,5}.*`), so a GPSLayer state_dict slice loads into it.
* `GPSBatch` -- the loader's collation (GraphGPS/graphgps/loader/batch.py:25-160): the rules of the root `batch.py` plus one: the
  per-graph `attn_bias` (flattened all-pairs shortest-path matrix from `create_subgraphs_gps`) is NOT collated, the batch carries
  `attn_bias = None` (:73,132-133).
"""
import copy

import torch
from torch.nn import Dropout, ELU, Sequential

from . import ops
from .batch import Batch
from .ops import BatchNorm1d as BN
from .ops import Linear


class ESCEdgeEncoding(torch.nn.Module):
    def __init__(self, dim_h, dropout):
        super(ESCEdgeEncoding, self).__init__()
        z_in = 1800
        hidden = dim_h
        self.z_initial = torch.nn.Embedding(z_in, hidden)
        self.z_embedding = Sequential(Dropout(dropout), BN(hidden), ELU(), Linear(hidden, hidden), Dropout(dropout), BN(hidden), ELU())

    def forward(self, batch):
        if hasattr(batch, 'pos_index'):
            index = ops.graph_index(batch)
            z_emb = ops.bag_embed_data(self.z_initial.weight, batch, index)
            batch.edge_attr = batch.edge_attr + self.z_embedding(z_emb)
        return batch


class GPSBatch(Batch):
    attn_bias = None                # loader/batch.py:132-133: the key survives collation with the value None

    @staticmethod
    def from_data_list(data_list, follow_batch=[]):
        stripped = []
        for d in data_list:
            if 'attn_bias' in d.keys:                      # shallow copy without the key: the caller's object is left intact
                c = copy.copy(d)
                object.__setattr__(c, '_store', {k: v for k, v in d._store.items() if k != 'attn_bias'})
                d = c
            stripped.append(d)
        out = Batch.from_data_list(stripped, follow_batch)
        out.__class__ = GPSBatch
        return out
