"""Inputs of the collation (B1) and `Distance` (D1) parity cases, shared by the fixture generator
tests/golden/make_golden_batch.py (which feeds them to the unmodified reference) and by the tests (which feed them to the
product).  No reference import here: the tests must run where /root/reference does not exist."""
import torch

from tests import model_util as MU

# (name, config, start, count, variant): variant 0 = plain; 1 = every third graph lacks edge_attr (keys missing in some graphs);
# 2 = follow_batch=['x', 'pos_enc']; 3 = original_edge_index / original_num_nodes keys present (run_qm9-style)
BATCH_CASES = [('cfg1', 1, 300, 9, 0), ('cfg2', 2, 300, 9, 0), ('cfg3', 3, 300, 5, 0), ('cfg4', 4, 300, 9, 0),
               ('cfg2_missing', 2, 320, 7, 1), ('cfg1_follow', 1, 330, 6, 2), ('cfg2_original', 2, 340, 5, 3)]


def batch_inputs(cls, config, start, count, variant):
    """The list of Data objects of one collation case (shared with tests/test_batch_cpu.py)."""
    out = []
    for i, g in enumerate(MU.graph_dicts(config, start, count)):
        d = cls(x=g['x'], edge_index=g['edge_index'], edge_attr=g.get('edge_attr'), y=g['y'].view(-1) if g['y'].dim() == 0 else g['y'],
                pos_enc=g['pos_enc'], pos_index=g['pos_index'], pos_batch=g['pos_batch'])
        if variant == 1 and i % 3 == 1:
            d.edge_attr = None
        if variant == 3:
            d.original_edge_index = g['edge_index'].flip(0).clone()
            d.original_num_nodes = int(g['num_nodes'])
        out.append(d)
    return out


def follow(variant):
    return ['x', 'pos_enc'] if variant == 2 else []


# (name, kwargs, has edge_attr, 1-D edge_attr, original_* keys)
DIST_CASES = [('default', dict(), True, False, False), ('nonorm', dict(norm=False), True, False, False),
              ('squared', dict(squared=True), True, False, False), ('relpos', dict(relative_pos=True), True, False, False),
              ('maxval', dict(max_value=3.0), True, False, False), ('nocat', dict(cat=False), True, False, False),
              ('noattr', dict(), False, False, False), ('attr1d', dict(), True, True, False),
              ('original', dict(squared=True), True, False, True), ('original_nonorm', dict(norm=False), True, True, True)]


def dist_inputs(seed, has_attr, one_d, original):
    g = torch.Generator().manual_seed(seed)
    n, e = 37, 150
    d = dict(pos=torch.randn(n, 3, generator=g), edge_index=torch.randint(0, n, (2, e), generator=g))
    if has_attr:
        d['edge_attr'] = torch.randn(e, generator=g) if one_d else torch.randn(e, 4, generator=g)
    if original:
        d['original_pos'] = torch.randn(n + 5, 3, generator=g)
        d['original_edge_index'] = torch.randint(0, n + 5, (2, e - 20), generator=g)
        d['original_edge_attr'] = torch.randn(e - 20, 2, generator=g)
    return d
