"""Loader for tests/golden/*.npz (layout documented in tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
FILES = ('kat1', 'edge_cases', 'sr25', 'cfg1', 'cfg2', 'cfg3', 'cfg4', 'cfg5')


class Case(object):
    def __init__(self, name, z):
        self.name = name
        self.ei = z[name + '/ei'].astype(np.int64)
        self.n, self.h, rd, sl = (int(v) for v in z[name + '/meta'])
        self.use_rd, self.self_loop = bool(rd), bool(sl)
        self.eo = z[name + '/eo'].astype(np.int64)
        self.nnz = z[name + '/nnz'].astype(np.int64)
        self.idx = z[name + '/idx'].astype(np.int64)
        self.cnt = z[name + '/cnt'].astype(np.int64)

    @property
    def pos_batch(self):
        return np.repeat(np.arange(self.eo.shape[1], dtype=np.int64), self.nnz)

    def dense(self):
        return dense(self.eo.shape[1], self.pos_batch, self.idx, self.cnt, self.use_rd)


def dense(num_edges, pos_batch, pos_index, pos_enc, use_rd):
    out = np.zeros((num_edges, 1800 if use_rd else 1700), dtype=np.int64)
    out[pos_batch, pos_index] = pos_enc
    return out


def load(fname):
    z = np.load(os.path.join(GOLDEN, fname + '.npz'))
    names = sorted({k.split('/')[0] for k in z.files})
    return [Case(nm, z) for nm in names]


def all_cases():
    for f in FILES:
        for c in load(f):
            yield c


def check_against_case(case, eo, pos_enc, pos_index, pos_batch, rd_exact=False):
    """Integer blocks bit-exact; rd block [400,500) only checked for its per-edge total unless rd_exact."""
    assert np.array_equal(np.asarray(eo), case.eo), case.name
    got = dense(case.eo.shape[1], np.asarray(pos_batch), np.asarray(pos_index), np.asarray(pos_enc), case.use_rd)
    ref = case.dense()
    if case.use_rd and not rd_exact:
        keep = np.ones(ref.shape[1], dtype=bool)
        keep[400:500] = False
        assert np.array_equal(got[:, keep], ref[:, keep]), case.name
        assert np.array_equal(got[:, 400:500].sum(1), ref[:, 400:500].sum(1)), case.name
    else:
        assert np.array_equal(got, ref), case.name
        # sparse layout itself: ascending index inside each edge, edges in order
        assert np.array_equal(np.asarray(pos_index), case.idx) and np.array_equal(np.asarray(pos_batch), case.pos_batch)
