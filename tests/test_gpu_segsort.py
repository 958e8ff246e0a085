"""The segmented hit sort (csrc/hitsort.cu: partition by the top key bits, per-bin sort in shared
memory) against the radix passes and the oracle: the hit list of motif_both_points.cpp:224-245 --
per query, per first table, ascending db id -- must come out byte for byte the same whichever
path orders it, in both output formats, including the range path of bins larger than the
shared-memory buffer, both per-bin sorts (buckets, radix passes) and the hand-back to the radix passes."""
import numpy as np
import pytest

import hsearch_b200 as hb
from tests.util import hits_as_tuples, planted_queries, random_codes

pytestmark = pytest.mark.gpu

MODES = {
    "radix": {"HS_SEGSORT": "0"},
    "seg": {"HS_SEGSORT": "1", "HS_SEGSORT_MIN": "0"},
    "seg_radix": {"HS_SEGSORT": "1", "HS_SEGSORT_MIN": "0", "HS_SEGSORT_RADIX": "1"},   # per-bin sort on the radix passes, not the buckets
    "seg_ranges": {"HS_SEGSORT": "1", "HS_SEGSORT_MIN": "0", "HS_SEGSORT_BUF": "48"},
    "seg_handback": {"HS_SEGSORT": "1", "HS_SEGSORT_MIN": "0", "HS_SEGSORT_BUF": "0"},   # no bin fits: every list is handed back
}


def set_mode(monkeypatch, name):
    for k in ("HS_SEGSORT", "HS_SEGSORT_MIN", "HS_SEGSORT_BUF", "HS_SEGSORT_RADIX"):
        monkeypatch.delenv(k, raising=False)
    for k, v in MODES[name].items():
        monkeypatch.setenv(k, v)   # read once, by hs_create


def make(length, K, L, W, R, seed=12345, **kw):
    h = hb.HSearch(length, K, L, W, R, **kw)
    a, b = h.seed_projection(seed)
    return h, a, b


@pytest.mark.parametrize("nq,id_base", [(300, 0), (2600, 0), (300, 5_000_000_000)])
def test_segsort_search_equals_radix_and_oracle(oracle, monkeypatch, nq, id_base):
    """hs_search_codes (hs_hit records; >= 2048 queries: the query-block pipeline) and
    hs_search_points_compact (CSR) under every ordering path."""
    length, K, L, W, R = 10, 4, 4, 50.0, 30.0
    codes = random_codes(60000, length, seed=41)
    qcodes = planted_queries(codes, nq, seed=42, frac=0.5)
    tab = oracle.coordinates(True)
    qpts = oracle.embed(qcodes, tab)
    res = {}
    for mode in MODES:
        set_mode(monkeypatch, mode)
        h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
        h.load_fragments(codes, id_base=id_base)
        h.build_index()
        plain = h.search_codes(qcodes)
        st = h.stats()
        compact = h.search_points_compact(qpts, expand=True)
        st2 = h.stats()
        res[mode] = (plain, compact)
        if mode == "radix":
            assert st.segsort_lists == 0 and st.segsort_fallbacks == 0
        elif mode == "seg_handback":
            assert st.segsort_fallbacks > 0 and st2.segsort_fallbacks > 0
        else:
            assert st.segsort_lists > 0 and st.segsort_fallbacks == 0
            assert st2.segsort_lists > 0 and st2.segsort_fallbacks == 0
        h.close()
    base = res["radix"]
    assert len(base[0]) > 500
    for mode, (plain, compact) in res.items():
        assert np.array_equal(plain, base[0]), mode
        assert np.array_equal(compact, base[1]), mode
        assert np.array_equal(compact, plain), mode
    want, _, _ = oracle.search(oracle.embed(codes, tab), qpts[:200], a, b, W, R)
    got = base[0][base[0]["query"] < 200].copy()
    got["db_id"] -= id_base
    assert hits_as_tuples(got) == hits_as_tuples(want)


def test_segsort_large_bins_brute_force(oracle, monkeypatch):
    """Brute force with a wide threshold, 9000 dense queries of which twelve are DB-like points and the
    rest lie far from every fragment: the bins are the queries, twelve of them hold tens of thousands
    of hits -- several ranking steps per radix pass, more keys than one shared-memory buffer (the range
    path without any test hook).  All paths must agree, and the list is in (query, db id) order."""
    length, R = 10, 44.0
    codes = random_codes(200000, length, seed=51)
    tab = oracle.coordinates(True)
    nq = 9000
    qpts = np.full((nq, 8 * length), 20.0)            # |q|^2 = 32000 (representable in the FP16 filter), > 90 from any fragment
    near = np.arange(12) * 700 + 5
    qpts[near] = oracle.embed(planted_queries(codes, 12, seed=52, frac=0.5), tab)
    res, used = {}, {}
    for mode in ("radix", "seg", "seg_radix", "seg_ranges"):
        set_mode(monkeypatch, mode)
        h, a, b = make(length, 4, 4, 50.0, R, flags=hb.HS_FLAG_SORT_HITS)
        h.load_fragments(codes)
        hits = h.bruteforce_points(qpts, cap=1 << 20)
        st = h.stats()
        used[mode] = (int(st.segsort_lists), int(st.segsort_fallbacks), len(hits))
        res[mode] = hits
        h.close()
    print("segsort (lists, hand-backs, hits) per mode:", used)
    base = res["radix"]
    assert set(np.unique(base["query"]).tolist()) == set(near.tolist())
    per_query = np.bincount(base["query"], minlength=nq)
    assert per_query.max() > 30000, per_query.max()    # larger than one shared-memory buffer (22528 keys)
    key = base["query"].astype(np.int64) * (1 << 32) + base["db_id"].astype(np.int64)
    assert np.all(np.diff(key) > 0)
    for mode in ("seg", "seg_radix", "seg_ranges"):
        assert np.array_equal(res[mode], base), mode
    assert used["radix"][:2] == (0, 0)
    assert used["seg"][:2] == (1, 0) and used["seg_radix"][:2] == (1, 0), used   # bins of > 100 k keys: taken in ranges
    assert used["seg_ranges"][:2] == (0, 1), used    # a 48-key buffer cannot hold one value of the top 8 key bits: handed back


def test_segsort_empty_and_tiny_lists(monkeypatch):
    """No hit at all, one hit, a handful of hits."""
    length = 10
    codes = random_codes(3000, length, seed=61)
    for mode in ("radix", "seg"):
        set_mode(monkeypatch, mode)
        h, a, b = make(length, 4, 4, 50.0, 1e-6, flags=hb.HS_FLAG_SORT_HITS)
        h.load_fragments(codes)
        h.build_index()
        q = codes[[5, 17, 17, 2999]].copy()
        hits = h.search_codes(q)          # each query finds exactly itself (and duplicates of itself)
        assert len(hits) >= 4 and set(hits["query"].tolist()) == {0, 1, 2, 3}
        assert np.all(hits["dist2"] == 0.0)
        q2 = (codes[:3] + 1) % 20         # shifted strings: nothing within 1e-6
        none = h.search_codes(q2.astype(np.uint8))
        assert len(none) == 0
        h.close()
