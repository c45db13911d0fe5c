"""Inputs of the GraphGPS-twin parity cases, shared by tests/golden/make_golden_gps.py (reference side) and
tests/test_gps_*.py (product side).  No reference import here."""
import torch

from tests import model_util as MU

GPS_BATCH = (2, 500, 7)            # config, start, count: ZINC-shaped graphs with an attn_bias key
GPS_INJECT = (64, 2, 520, 9)       # dim_h, config, start, count


def gps_graphs(cls, config, start, count):
    """Per-graph Data objects as loader/utils_escgnn.create_subgraphs returns them: the encodings plus a flattened [n*n] attn_bias."""
    out = []
    for g in MU.graph_dicts(config, start, count):
        n = int(g['num_nodes'])
        gen = torch.Generator().manual_seed(n)
        d = cls(x=g['x'], edge_index=g['edge_index'], edge_attr=g.get('edge_attr'), y=g['y'].view(-1) if g['y'].dim() == 0 else g['y'],
                pos_enc=g['pos_enc'], pos_index=g['pos_index'], pos_batch=g['pos_batch'],
                attn_bias=torch.randint(0, 9, (n * n, ), generator=gen))
        out.append(d)
    return out


def inject_batch(config, start, count, dim_h):
    """A collated batch whose edge_attr already is a [E, dim_h] float embedding (what the GraphGPS edge encoder produced)."""
    b = MU.ref_batch(config, start, count)
    gen = torch.Generator().manual_seed(99)
    b.edge_attr = torch.randn(b.edge_index.size(1), dim_h, generator=gen)
    return b
