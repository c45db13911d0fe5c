"""CPU: both oracles (numpy literal restatement, fast C restatement) against the reference-generated fixtures."""
import numpy as np
import pytest

from oracle import c_oracle, encode_ref
from tests import golden_util as G


@pytest.mark.parametrize('fname', G.FILES)
def test_c_oracle_matches_reference_fixtures(fname):
    for c in G.load(fname):
        out = c_oracle.encode_graph(c.ei, c.n, c.h, c.use_rd, c.self_loop)
        G.check_against_case(c, *out)


@pytest.mark.parametrize('fname', ('kat1', 'edge_cases', 'cfg2', 'cfg4'))
def test_numpy_oracle_matches_reference_fixtures(fname):
    for c in G.load(fname)[:24]:
        out = encode_ref.encode_graph(c.ei, c.n, c.h, c.use_rd, c.self_loop)
        G.check_against_case(c, *out)


def test_numpy_and_c_oracle_agree_including_rd():
    for c in G.load('cfg1')[:6] + G.load('cfg2')[:10] + [x for x in G.load('edge_cases') if x.use_rd]:
        a = encode_ref.encode_graph(c.ei, c.n, c.h, c.use_rd, c.self_loop)
        b = c_oracle.encode_graph(c.ei, c.n, c.h, c.use_rd, c.self_loop)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), c.name


def test_rd_policy_E5_disagreements_sit_on_integer_resistances():
    """wherever the literal (float32 LAPACK) reference and the fp64 oracle bin a node
    differently, the exact resistance distance is within 1e-4 of an integer."""
    checked = moved = 0
    for c in G.load('cfg2')[:12] + G.load('cfg1')[:4]:
        rd_raw = []
        _, pe, pi, pb = encode_ref.encode_graph(c.ei, c.n, c.h, True, c.self_loop, rd_out=rd_raw)
        got = G.dense(c.eo.shape[1], pb, pi, pe, True)[:, 400:500]
        ref = c.dense()[:, 400:500]
        for e in range(ref.shape[0]):
            checked += 1
            if np.array_equal(got[e], ref[e]):
                continue
            moved += 1
            r = rd_raw[e]
            near = np.abs(r - np.round(r)) < 1e-4
            # bins that differ must be explained by near-integer nodes moving between k-1 and k
            diff = got[e] - ref[e]
            assert diff.sum() == 0
            assert np.abs(diff).sum() <= 2 * near.sum(), (c.name, e)
            for k in np.nonzero(diff)[0]:
                assert np.any(near & ((np.round(r) == k) | (np.round(r) == k + 1))), (c.name, e, k)
    assert checked > 500 and moved > 0


def test_kat1_known_answers():
    """SURVEY section 4 KAT-1 (values printed by the unmodified reference)."""
    und = [(0, 1), (1, 2), (2, 3), (1, 4), (2, 4)]
    ei = np.array(und + [(b, a) for a, b in und]).T
    eo, pe, pi, pb = c_oracle.encode_graph(ei, 5, 2, False, False)
    assert eo.shape[1] == 10 and pe.size == 182
    e0 = dict(zip(pi[pb == 0].tolist(), pe[pb == 0].tolist()))
    assert e0 == {1: 2, 2: 1, 3: 2, 200: 1, 201: 1, 202: 2, 203: 1, 300: 1, 301: 3, 302: 1, 442: 1, 617: 1,
                  629: 2, 874: 2, 881: 2, 888: 1, 1133: 1}
    eo, pe, pi, pb = c_oracle.encode_graph(ei, 5, 2, False, True)
    assert eo.shape[1] == 15 and pe.size == 257
    e10 = dict(zip(pi[pb == 10].tolist(), pe[pb == 10].tolist()))
    assert e10 == {0: 1, 2: 1, 3: 2, 4: 1, 200: 2, 201: 1, 202: 2, 300: 2, 301: 1, 302: 2, 407: 1, 652: 1,
                   666: 2, 911: 2, 918: 2}
    assert c_oracle.encode_graph(ei, 5, 2, True, False)[1].size == 204
    assert c_oracle.encode_graph(ei, 5, 2, True, True)[1].size == 284


def test_sr25_kat2_counts():
    """SURVEY KAT-2: distinct per-graph multisets h=1 -> 15, h=2 -> 9, h=3 -> 9; graph #15 nnz and sum."""
    cases = G.load('sr25')
    for h, n_distinct, nnz15, sum15 in ((1, 15, 8370, 61914), (2, 9, 10250, 121950), (3, 9, 10250, 121950)):
        sigs = set()
        for c in [x for x in cases if x.h == h]:
            out = c_oracle.encode_graph(c.ei, c.n, c.h, False, True)
            d = G.dense(out[0].shape[1], out[3], out[2], out[1], False)
            sigs.add(tuple(sorted(map(bytes, d))))
            if c.name.startswith('sr25_g14'):
                assert out[1].size == nnz15 and out[1].sum() == sum15
        assert len(sigs) == n_distinct


def test_oracle_error_behaviour():
    hub = np.array([[0] * 210 + list(range(1, 211)), list(range(1, 211)) + [0] * 210])
    with pytest.raises(ValueError):
        c_oracle.encode_graph(hub, 211, 1)           # degree >= 200 (reference: F.one_hot raises)
    with pytest.raises(ValueError):
        c_oracle.encode_graph(hub[:, :4], 211, 5)    # h >= 5
    with pytest.raises(ValueError):
        c_oracle.encode_graph(np.array([[0, 1], [1, 2]]), 3, 2, use_rd=True)   # rd on a directed multiset


def test_all_pairs_spd_oracle_matches_networkx_fixture():
    """N4: the attn_bias restatement against vectors produced by networkx (the routine the reference calls,
    GraphGPS/graphgps/loader/utils_escgnn.py:29-38): connected, disconnected, self-loop, one-directional edge, long path."""
    import os
    import numpy as np
    from oracle import encode_ref
    fix = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'spd.npz'))
    names = sorted({k.split('/')[0] for k in fix.files})
    assert len(names) == 8
    for nm in names:
        got = encode_ref.all_pairs_spd(fix[nm + '/edge_index'], int(fix[nm + '/n'][0]))
        assert np.array_equal(got, fix[nm + '/attn_bias']), nm
