"""Minimal `Data` container with the PyG 2.0.4 behaviours the reference transform / collation rely on.

The reference passes `torch_geometric.data.Data` objects (utils_edge_efficient.py:29, batch.py:10); PyG is not
installed here, and the drop-in transform only needs this protocol: positional ctor `(x, edge_index, edge_attr,
y, pos, **kw)`, attribute access that yields None for a missing core key and AttributeError for any other missing
key, `key in data`, `data[key]`, `.keys`, `.num_nodes`, `__cat_dim__`, `__inc__`.  Any class honouring it works
with `create_subgraphs` (a real PyG `Data` does).
"""
import torch

_CORE = ('x', 'edge_index', 'edge_attr', 'y', 'pos')


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        object.__setattr__(self, '_store', {})
        for k, v in zip(_CORE, (x, edge_index, edge_attr, y, pos)):
            if v is not None:
                self._store[k] = v
        for k, v in kwargs.items():
            if v is not None:
                self._store[k] = v

    def __getattr__(self, key):
        if key.startswith('__') and key.endswith('__'):
            raise AttributeError(key)
        store = object.__getattribute__(self, '_store')
        if key in store:
            return store[key]
        if key in _CORE:
            return None
        raise AttributeError("'%s' object has no attribute '%s'" % (self.__class__.__name__, key))

    def __setattr__(self, key, value):
        if key.startswith('__') and key.endswith('__'):
            object.__setattr__(self, key, value)
        elif key == 'num_nodes':
            self._store['__num_nodes__'] = value
        elif value is None:
            self._store.pop(key, None)
        else:
            self._store[key] = value

    def __delattr__(self, key):
        self._store.pop(key, None)

    def __getitem__(self, key):
        return self._store.get(key, None)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.keys

    @property
    def keys(self):
        return [k for k in self._store if not (k.startswith('__') and k.endswith('__'))]

    def __iter__(self):
        for k in sorted(self.keys):
            yield k, self._store[k]

    def __len__(self):
        return len(self.keys)

    @property
    def num_nodes(self):
        if '__num_nodes__' in self._store:
            return self._store['__num_nodes__']
        if 'x' in self._store:
            return self._store['x'].size(0)
        if 'pos' in self._store:
            return self._store['pos'].size(0)
        if 'edge_index' in self._store and self._store['edge_index'].numel() > 0:
            return int(self._store['edge_index'].max()) + 1
        return None

    @property
    def num_edges(self):
        ei = self._store.get('edge_index')
        return None if ei is None else ei.size(1)

    def __cat_dim__(self, key, value):
        return -1 if 'index' in key else 0

    def __inc__(self, key, value):
        return self.num_nodes if 'index' in key else 0

    def to(self, device, *args, **kwargs):
        for key, v in list(self._store.items()):
            if torch.is_tensor(v):
                self._store[key] = v.to(device, *args, **kwargs)
        return self

    def contiguous(self):
        for key, v in list(self._store.items()):
            if torch.is_tensor(v):
                self._store[key] = v.contiguous()
        return self

    def __repr__(self):
        return '%s(%s)' % (self.__class__.__name__, ', '.join(
            '%s=%s' % (k, list(v.shape) if torch.is_tensor(v) else v) for k, v in self))
