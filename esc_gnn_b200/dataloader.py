"""DataLoader that collates with `esc_gnn_b200.batch.Batch` (mirrors /root/reference/dataloader.py:11-48)."""
import torch.utils.data

from .batch import Batch
from .data import Data


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, follow_batch=[], **kwargs):
        def collate(batch):
            elem = batch[0]
            if isinstance(elem, Data) or hasattr(elem, 'edge_index'):
                return Batch.from_data_list(batch, follow_batch)
            if isinstance(elem, float):
                return torch.tensor(batch, dtype=torch.float)
            if isinstance(elem, int):
                return torch.tensor(batch)
            if isinstance(elem, (str, bytes)):
                return batch
            if isinstance(elem, dict):
                return {key: collate([d[key] for d in batch]) for key in elem}
            if isinstance(elem, (list, tuple)):
                return [collate(s) for s in zip(*batch)]
            raise TypeError('DataLoader found invalid type: {}'.format(type(elem)))

        super(DataLoader, self).__init__(dataset, batch_size, shuffle, collate_fn=collate, **kwargs)
