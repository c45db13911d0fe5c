"""`Distance` edge transform with the reference's constructor and semantics (/root/reference/distance.py:5-65),
computed by the sm_100a `edge_distance` kernel when the data sits on the GPU."""
import torch

from . import ops_distance


class Distance(object):
    def __init__(self, norm=True, max_value=None, cat=True, relative_pos=False, squared=False):
        self.norm = norm
        self.max = max_value
        self.cat = cat
        self.relative_pos = relative_pos
        self.squared = squared

    def _one(self, edge_index, pos, pseudo):
        dist, rel = ops_distance.edge_distance(pos, edge_index, self.squared, self.norm, self.max)
        if pseudo is not None and self.cat:
            pseudo = pseudo.view(-1, 1) if pseudo.dim() == 1 else pseudo
            out = torch.cat([pseudo, dist.type_as(pseudo)], dim=-1)
        else:
            out = dist
        return out, rel

    def __call__(self, data):
        if type(data) == dict:
            return {key: self.__call__(d) for key, d in data.items()}
        data.edge_attr, rel = self._one(data.edge_index, data.pos, data.edge_attr)
        if self.relative_pos:
            data.edge_attr = torch.cat([data.edge_attr, rel], dim=-1)
        if 'original_edge_index' in data:        # distance.py:49-63 (always un-squared there)
            saved = self.squared
            self.squared = False
            data.original_edge_attr, _ = self._one(data.original_edge_index, data.original_pos,
                                                   data.original_edge_attr)
            self.squared = saved
        return data

    def __repr__(self):
        return '{}(norm={}, max_value={})'.format(self.__class__.__name__, self.norm, self.max)
