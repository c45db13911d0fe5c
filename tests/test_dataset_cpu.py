"""CPU: the processed-dataset store (`collate` / `separate`, the InMemoryDataset `(data, slices)` layout the reference's
process() methods save -- GraphCountDataset.py:118-119, dataset_zinc.py:87-88) round-trips every graph."""
import torch

from esc_gnn_b200 import dataset
from esc_gnn_b200.data import Data
from tests import model_util as MU


def _data(g):
    return Data(x=g['x'], edge_index=g['edge_index'], edge_attr=g.get('edge_attr'), y=g['y'], pos_enc=g['pos_enc'],
                pos_index=g['pos_index'], pos_batch=g['pos_batch'])


def test_collate_separate_round_trip():
    for config in (1, 2, 4):
        ds = [_data(g) for g in MU.graph_dicts(config, 40, 7)]
        big, slices = dataset.collate(ds)
        assert big.edge_index.size(1) == sum(d.edge_index.size(1) for d in ds)          # concatenated along the last dim
        assert int(big.edge_index.max()) < max(d.num_nodes for d in ds)                 # and NOT incremented
        for i, want in enumerate(ds):
            got = dataset.separate(big, slices, i)
            assert sorted(got.keys) == sorted(want.keys)
            for k in want.keys:
                assert torch.equal(got[k].reshape(-1), want[k].reshape(-1)) and got[k].dtype == want[k].dtype, (config, i, k)
            assert got.edge_index.shape == want.edge_index.shape and got.num_nodes == want.num_nodes


def test_store_keeps_isolated_tail_nodes():
    a = Data(x=None, edge_index=torch.tensor([[0, 1], [1, 0]]), y=torch.tensor([1.0]))
    a.num_nodes = 5                                   # nodes 2..4 isolated: edge_index.max()+1 would lose them
    big, slices = dataset.collate([a, a])
    assert dataset.separate(big, slices, 1).num_nodes == 5
