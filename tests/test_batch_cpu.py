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


# ---- pinned to the reference: fixtures produced by the UNMODIFIED /root/reference/batch.py (tests/golden/make_golden_batch.py)
def _fixture():
    import os
    import numpy as np
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'batch.npz'))


def test_from_data_list_matches_reference_fixtures():
    import numpy as np
    from tests import batch_cases as MG
    fix = _fixture()
    names = sorted({k.split('/')[1] for k in fix.files if k.startswith('batch/')})
    assert len(names) >= 7
    for nm in names:
        config, start, count, variant = [int(v) for v in fix['batch/%s/meta' % nm]]
        got = Batch.from_data_list(MG.batch_inputs(Data, config, start, count, variant), follow_batch=MG.follow(variant))
        keys = sorted(k.split('/')[2] for k in fix.files if k.startswith('batch/%s/' % nm) and k.split('/')[2] not in ('meta', 'num_graphs'))
        assert sorted(k for k in got.keys if torch.is_tensor(got[k])) == keys, (nm, sorted(got.keys), keys)
        for k in keys:
            want = fix['batch/%s/%s' % (nm, k)]
            assert np.array_equal(got[k].numpy(), want) and got[k].numpy().dtype == want.dtype, (nm, k)
        assert got.num_graphs == int(fix['batch/%s/num_graphs' % nm][0])


def test_oracle_collate_matches_reference_fixtures():
    """The oracle restatement used by the model tests (oracle/model_ref.collate) is itself pinned to the reference output."""
    import numpy as np
    fix = _fixture()
    for nm in ('cfg1', 'cfg2', 'cfg3', 'cfg4'):
        config, start, count, _ = [int(v) for v in fix['batch/%s/meta' % nm]]
        want = model_ref.collate(MU.graph_dicts(config, start, count))
        for k in ('x', 'edge_index', 'pos_enc', 'pos_index', 'pos_batch', 'batch'):
            assert np.array_equal(getattr(want, k).numpy(), fix['batch/%s/%s' % (nm, k)]), (nm, k)
        if want.edge_attr is not None:
            assert np.array_equal(want.edge_attr.numpy(), fix['batch/%s/edge_attr' % nm])


def test_oracle_distance_matches_reference_fixtures():
    import numpy as np
    fix = _fixture()
    d = {k.split('/', 2)[2]: torch.as_tensor(fix[k]) for k in fix.files if k.startswith('dist/default/')}
    got = model_ref.distance_transform(d['in_edge_index'], d['in_pos'], d['in_edge_attr'])
    assert np.array_equal(got.numpy(), fix['dist/default/edge_attr'])
