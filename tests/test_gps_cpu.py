"""CPU: the GraphGPS loader collation (product `GPSBatch`) against the fixture produced by the unmodified
GraphGPS/graphgps/loader/batch.py (tests/golden/make_golden_gps.py)."""
import os

import numpy as np
import torch

from esc_gnn_b200.data import Data
from esc_gnn_b200.gps import GPSBatch
from tests.gps_cases import GPS_BATCH, gps_graphs

FIX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'gps.npz'))


def test_gps_batch_matches_reference_loader_batch():
    graphs = gps_graphs(Data, *GPS_BATCH)
    keep = [g.attn_bias.clone() for g in graphs]
    b = GPSBatch.from_data_list(graphs)
    keys = [k[len('gps/batch/'):] for k in FIX.files if k.startswith('gps/batch/') and not k.endswith(('has_attn_bias_key', 'attn_bias_is_none'))]
    assert keys
    for k in keys:
        assert torch.equal(b[k], torch.from_numpy(FIX['gps/batch/' + k])), k
    assert int(FIX['gps/batch/attn_bias_is_none'][0]) == 1 and b.attn_bias is None        # loader/batch.py:132-133
    assert b.num_graphs == GPS_BATCH[2]
    for g, a in zip(graphs, keep):                                                          # the per-graph objects are left intact
        assert torch.equal(g.attn_bias, a)
