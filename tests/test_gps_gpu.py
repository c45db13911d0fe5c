"""GPU: the GPSLayer edge-feature injection (product `ESCEdgeEncoding` on the sm_100a kernels) against the fixture produced by
the statements lifted from the unmodified GraphGPS/graphgps/layer/gps_layer.py:169-188 (tests/golden/make_golden_gps.py)."""
import os

import numpy as np
import pytest
import torch

from tests import model_util as MU
from tests.gps_cases import GPS_INJECT, inject_batch

pytestmark = pytest.mark.gpu
FIX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'gps.npz'))


def test_edge_encoding_injection_matches_reference_statements():
    from esc_gnn_b200.gps import ESCEdgeEncoding
    torch.backends.cuda.matmul.allow_tf32 = False
    dim_h, config, start, count = GPS_INJECT
    m = ESCEdgeEncoding(dim_h, 0.0).cuda()
    sd = MU.det_state(m.state_dict(), seed=4321)
    m.load_state_dict({k: v.cuda() for k, v in sd.items()})
    m.train()
    b = inject_batch(config, start, count, dim_h)
    for k, v in list(b.__dict__.items()):
        if torch.is_tensor(v):
            setattr(b, k, v.cuda())
    out = m(b).edge_attr
    want = torch.from_numpy(FIX['gps/inject/edge_attr_out']).cuda()
    assert (out - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())
    (out * torch.linspace(-1, 1, out.numel(), device='cuda').view_as(out)).sum().backward()
    grads = dict(m.named_parameters())
    for k, w in zip(FIX['gps/inject/grad_keys'], FIX['gps/inject/grad_digest']):
        got = MU.grad_digest(grads[str(k)].grad)
        if str(k) == 'z_embedding.3.bias':
            continue                                   # a Linear bias feeding BatchNorm: mathematically zero gradient (rounding noise)
        assert abs(got[3] - w[3]) <= 2e-3 * w[3] + 1e-6, (str(k), got[3], w[3])
