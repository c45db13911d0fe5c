"""CPU: collation rules of esc_gnn_b200.batch.Batch against the oracle restatement of reference batch.py:25-149."""
import torch

from esc_gnn_b200.batch import Batch
from esc_gnn_b200.data import Data
from esc_gnn_b200.dataloader import DataLoader
from oracle import model_ref
from tests import model_util as MU


def to_data(d):
    out = Data(x=d['x'], edge_index=d['edge_index'], edge_attr=d.get('edge_attr'), y=d['y'], pos_enc=d['pos_enc'],
               pos_index=d['pos_index'], pos_batch=d['pos_batch'])
    return out


def test_from_data_list_matches_reference_rules():
    for config in (1, 2, 4):
        gs = MU.graph_dicts(config, 300, 9)
        want = model_ref.collate(gs)
        got = Batch.from_data_list([to_data(g) for g in gs])
        for k in ('x', 'edge_index', 'y', 'pos_enc', 'pos_index', 'pos_batch', 'batch'):
            assert torch.equal(got[k] if k != 'batch' else got.batch, getattr(want, k)), (config, k)
        if want.edge_attr is not None:
            assert torch.equal(got.edge_attr, want.edge_attr)
        assert got.num_graphs == 9
        # pos_batch ends at the number of edges: the contract bag-embed relies on
        assert int(got.pos_batch[-1]) + 1 == got.edge_index.size(1)


def test_dataloader_yields_batches():
    gs = [to_data(g) for g in MU.graph_dicts(2, 0, 10)]
    loader = DataLoader(gs, batch_size=4, shuffle=False)
    sizes = [b.num_graphs for b in loader]
    assert sizes == [4, 4, 2]
