"""CPU: the sweep's order-independent checksum equals the C oracle's batch digest (so bench.py can check a sweep prefix against
the oracle without moving the records through Python loops), and the tiled chunk builder keeps graphs intact."""
import numpy as np

from esc_gnn_b200 import sweep, synth
from oracle import c_oracle


def test_digest_matches_c_oracle_batch_digest():
    for config, h, sl in ((5, 2, False), (5, 3, True), (2, 3, False)):
        src, dst, eptr, nptr = synth.make_batch_arrays(config, 40, 12)
        eo_ptr, pe, pi, pb, base = [0], [], [], [], 0
        for g in range(len(nptr) - 1):
            a, b = eptr[g], eptr[g + 1]
            r = c_oracle.encode_graph(np.stack([src[a:b], dst[a:b]]), int(nptr[g + 1] - nptr[g]), h, False, sl)
            pe.append(r[1]); pi.append(r[2]); pb.append(r[3] + base)
            base += r[0].shape[1]
            eo_ptr.append(base)
        got = sweep.digest(np.array(eo_ptr), base, np.concatenate(pe), np.concatenate(pi), np.concatenate(pb))
        want = c_oracle.encode_batch_digest(src, dst, eptr, nptr, h, False, sl, threads=2)
        assert got == want, (config, h, sl, got, want)


def test_tiled_chunk_repeats_whole_graphs():
    src, dst, eptr, nptr = sweep.tiled_chunk(5, 6, 16)
    assert len(eptr) == 17 and len(nptr) == 17 and eptr[-1] == len(src) == len(dst)
    s0, d0, e0, n0 = synth.make_batch_arrays(5, 0, 6)
    for g in range(16):
        k = g % 6
        assert np.array_equal(src[eptr[g]:eptr[g + 1]], s0[e0[k]:e0[k + 1]])
        assert nptr[g + 1] - nptr[g] == n0[k + 1] - n0[k]
