"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the
same seeded inputs.  Bit-exact for bucket ints, packed keys, bucket membership,
hit sets, hit order and FP64 distances."""
import numpy as np
import pytest

import hsearch_b200 as hb
from tests.util import hits_as_tuples, planted_families, planted_queries, random_codes

pytestmark = pytest.mark.gpu


def make(length, K, L, W, R, seed=12345, **kw):
    h = hb.HSearch(length, K, L, W, R, **kw)
    a, b = h.seed_projection(seed)
    return h, a, b


def table_for(oracle, variant):
    return oracle.coordinates(print6=(variant == hb.HS_TABLE_PRINT6))


# ---------------------------------------------------------------- K1: hash
@pytest.mark.parametrize("length,K,L,W", [(10, 4, 4, 50.0), (10, 4, 4, 20.0), (10, 4, 4, 4.0), (25, 4, 4, 50.0),
                                          (8, 2, 1, 10.0), (30, 8, 3, 50.0), (10, 16, 2, 50.0), (12, 3, 5, 7.5),
                                          (10, 16, 32, 10.0), (1, 1, 1, 1.0)])
@pytest.mark.parametrize("variant", [hb.HS_TABLE_PRINT6, hb.HS_TABLE_FULL])
def test_hash_buckets_and_keys(oracle, length, K, L, W, variant):
    n = 20000 if K * L <= 64 else 3000
    codes = random_codes(n, length, seed=1)
    h, a, b = make(length, K, L, W, 30.0, table_variant=variant, flags=hb.HS_FLAG_HASH_AUDIT)
    h.load_fragments(codes)
    got = h.hash(want_buckets=True)
    want = oracle.hash_codes(codes, table_for(oracle, variant), a, b, W)
    assert np.array_equal(got, want)
    st = h.stats()
    assert st.residual_flips == 0
    kw = st.key_words
    strings = oracle.key_strings(want)
    for l in range(L):
        keys = h.keys(l)
        uniq = {}
        for s in set(strings[:, l].tolist()):
            uniq[s] = hb.pack_key_string(s, kw)
        expect = np.stack([uniq[s] for s in strings[:, l]])
        assert np.array_equal(keys, expect)
    h.close()


def test_hash_projection_matches_oracle_rng(oracle):
    a, b = hb.generate_projection(12345, 80, 4, 4, 50.0)
    oa, ob = oracle.lsh_tables(12345, 80, 4, 4, 50.0)
    assert np.array_equal(a, oa) and np.array_equal(b, ob)


def test_hash_exact_mode_equals_fast_mode(oracle):
    codes = random_codes(50000, 10, seed=3)
    res = []
    for flags in (0, hb.HS_FLAG_HASH_EXACT):
        h, a, b = make(10, 4, 4, 50.0, 30.0, flags=flags)
        h.load_fragments(codes)
        res.append(h.hash(want_buckets=True))
        if flags == 0:
            st = h.stats()
            # the guard band is exercised but rare
            assert st.guard_hits < 0.01 * codes.shape[0] * 16
        h.close()
    assert np.array_equal(res[0], res[1])


def test_hash_guard_band_forced(oracle):
    """W tiny -> nearly every projection sits inside the guard band; the FP64
    re-evaluation must still reproduce the oracle."""
    codes = random_codes(5000, 10, seed=4)
    W = 1e-3
    h = hb.HSearch(10, 2, 2, W, 30.0)
    a, b = hb.generate_projection(7, 80, 2, 2, W)
    try:
        h.set_projection(a, b)
    except hb.HsError as e:  # key too wide for this W is a legitimate refusal
        assert e.code == -4
        return
    h.load_fragments(codes)
    got = h.hash(want_buckets=True)
    want = oracle.hash_codes(codes, oracle.coordinates(True), a, b, W)
    assert np.array_equal(got, want)
    assert h.stats().guard_hits > 0
    h.close()


# ---------------------------------------------------------------- K2: sort + buckets
@pytest.mark.parametrize("n", [1, 2, 255, 4096, 4097, 100003])
@pytest.mark.parametrize("K,W", [(4, 50.0), (4, 4.0), (16, 50.0)])
def test_index_buckets(oracle, n, K, W):
    L = 3
    codes = random_codes(n, 10, seed=n)
    h, a, b = make(10, K, L, W, 30.0)
    h.load_fragments(codes)
    h.build_index()
    strings = oracle.key_strings(oracle.hash_codes(codes, oracle.coordinates(True), a, b, W))
    sizes = h.table_sizes()
    for l in range(L):
        ids, starts = h.table(l)
        assert sorted(ids.tolist()) == list(range(n))
        groups = {}
        for i, s in enumerate(strings[:, l]):
            groups.setdefault(s, []).append(i)
        assert sizes[l] == len(groups)
        assert starts[0] == 0 and starts[-1] == n
        seen = set()
        for bidx in range(len(starts) - 1):
            members = ids[starts[bidx]:starts[bidx + 1]].tolist()
            s = strings[members[0], l]
            assert members == groups[s], "bucket members must be the ascending ids sharing the key string"
            seen.add(s)
        assert len(seen) == len(groups)
    h.close()


# ---------------------------------------------------------------- K3: search
@pytest.mark.parametrize("length,K,L,W,R", [(10, 4, 4, 50.0, 30.0), (10, 4, 4, 20.0, 30.0), (10, 2, 6, 30.0, 25.0),
                                            (25, 4, 4, 50.0, 60.0), (8, 4, 4, 50.0, 30.0), (30, 4, 2, 80.0, 70.0),
                                            (10, 16, 4, 200.0, 30.0)])
def test_search_hits_bit_exact(oracle, length, K, L, W, R):
    n, q = 30000, 300
    codes = random_codes(n, length, seed=11)
    qcodes = planted_queries(codes, q, seed=12)
    tab = oracle.coordinates(True)
    h, a, b = make(length, K, L, W, R)
    h.load_fragments(codes)
    h.build_index()
    got = h.search_codes(qcodes)
    want, ts, ncand = oracle.search(oracle.embed(codes, tab), oracle.embed(qcodes, tab), a, b, W, R, pred=0)
    assert np.array_equal(h.table_sizes(), ts)
    assert len(got) == len(want) and len(want) > 0
    assert hits_as_tuples(got) == hits_as_tuples(want)  # order, first table, id, FP64 distance
    h.close()


def test_search_dense_query_points(oracle):
    """Centres may be arbitrary real vectors (cluster means), not embeddings."""
    n, q, length = 20000, 200, 10
    codes = random_codes(n, length, seed=21)
    tab = oracle.coordinates(True)
    rng = np.random.default_rng(22)
    qpts = oracle.embed(planted_queries(codes, q, seed=23), tab) + rng.normal(0, 0.3, size=(q, 8 * length))
    h, a, b = make(length, 4, 4, 50.0, 30.0)
    h.load_fragments(codes)
    h.build_index()
    got = h.search_points(qpts)
    want, ts, _ = oracle.search(oracle.embed(codes, tab), qpts, a, b, 50.0, 30.0, pred=0)
    assert len(want) > 0
    assert hits_as_tuples(got) == hits_as_tuples(want)
    h.close()


def test_search_edge_cases(oracle):
    tab = oracle.coordinates(True)
    codes = random_codes(5000, 10, seed=31)
    h, a, b = make(10, 4, 4, 50.0, 30.0)
    h.load_fragments(codes)
    h.build_index()
    # no queries
    assert len(h.search_codes(np.zeros((0, 10), dtype=np.uint8))) == 0
    # a query far from everything: dense point way outside the embedding
    far = np.full((1, 80), 1e4)
    assert len(h.search_points(far)) == 0
    # duplicates in the DB and query == DB fragment (distance 0)
    got = h.search_codes(codes[:5])
    want, _, _ = oracle.search(oracle.embed(codes, tab), oracle.embed(codes[:5], tab), a, b, 50.0, 30.0)
    assert hits_as_tuples(got) == hits_as_tuples(want)
    assert any(t[3] == 0.0 for t in hits_as_tuples(got))
    # capacity error path: caller buffer too small -> retried inside the wrapper
    got2 = h.search_codes(codes[:50], cap=1)
    want2, _, _ = oracle.search(oracle.embed(codes, tab), oracle.embed(codes[:50], tab), a, b, 50.0, 30.0)
    assert hits_as_tuples(got2) == hits_as_tuples(want2)
    h.close()
    # empty database
    h, a, b = make(10, 4, 4, 50.0, 30.0)
    h.load_fragments(np.zeros((0, 10), dtype=np.uint8))
    h.build_index()
    assert len(h.search_codes(codes[:3])) == 0
    h.close()


def test_search_against_reference_itself(oracle, reference):
    """The reference's own Search() (compiled in place) on the same inputs."""
    n, q = 20000, 200
    codes = random_codes(n, 10, seed=41)
    qcodes = planted_queries(codes, q, seed=42)
    tab = oracle.coordinates(True)
    db, qp = oracle.embed(codes, tab), oracle.embed(qcodes, tab)
    for W in (20.0, 50.0):
        h, a, b = make(10, 4, 4, W, 30.0, seed=12345)
        h.load_fragments(codes)
        h.build_index()
        got = h.search_codes(qcodes)
        ref, printed, ts, _ = reference.search(db, qp, 4, 4, W, 30.0, 12345)
        assert np.array_equal(h.table_sizes(), ts)
        assert hits_as_tuples(got, with_table=False) == hits_as_tuples(ref, with_table=False)
        h.close()


# ---------------------------------------------------------------- K4: brute force
@pytest.mark.parametrize("length,R", [(10, 30.0), (25, 60.0), (8, 25.0)])
def test_bruteforce_points(oracle, length, R):
    n, q = 20000, 100
    codes = random_codes(n, length, seed=51)
    qcodes = planted_queries(codes, q, seed=52)
    tab = oracle.coordinates(True)
    h = hb.HSearch(length, 4, 4, 50.0, R, predicate=hb.HS_PRED_SQRT_LE_R)
    h.load_fragments(codes)
    got = h.bruteforce_codes(qcodes)
    want = oracle.bruteforce(oracle.embed(codes, tab), oracle.embed(qcodes, tab), R, pred=1)
    assert len(want) > 0
    assert hits_as_tuples(got, False) == hits_as_tuples(want, False)
    h.close()


@pytest.mark.parametrize("length,R", [(8, 30), (10, 40), (16, 80), (30, 200)])
def test_bruteforce_int_metric_all_pairs(oracle, length, R):
    n = 3000
    codes = planted_families(n, length, seed=61)
    h = hb.HSearch(length, 4, 4, 50.0, float(R), metric=hb.HS_METRIC_BLOSUM_INT)
    h.load_fragments(codes)
    got = h.bruteforce_codes(None, cap=1 << 22)
    want = oracle.bruteforce_int(codes, None, R)
    assert len(want) > 0
    assert hits_as_tuples(got, False) == hits_as_tuples(want, False)
    # and with explicit queries
    got = h.bruteforce_codes(codes[:64], cap=1 << 22)
    want = oracle.bruteforce_int(codes, codes[:64], R)
    assert hits_as_tuples(got, False) == hits_as_tuples(want, False)
    h.close()


@pytest.mark.parametrize("length,R", [(8, 24), (10, 34), (12, 40), (25, 90), (30, 120)])
def test_int_metric_on_pipelined_tensor_filter(oracle, monkeypatch, length, R):
    """The integer BLOSUM metric runs through the pipelined tensor filter with a contracting embedding
    (hs_get_blosum_filter_embedding): all-pairs, explicit-query brute force, LSH search and cluster
    must equal the one-hot tensor filter (HS_NO_MMA_INT) and the scalar filter, and the oracle."""
    n = 40000
    codes = np.concatenate([planted_families(6000, length, seed=161), random_codes(n - 6000, length, seed=162)])
    qcodes = planted_queries(codes[:6000], 300, seed=163)
    res = []
    for env, flags in (({}, 0), ({"HS_NO_MMA_INT": "1"}, 0), ({}, hb.HS_FLAG_SCALAR_FILTER)):
        monkeypatch.delenv("HS_NO_MMA_INT", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        h = hb.HSearch(length, 4, 4, 50.0, float(R), metric=hb.HS_METRIC_BLOSUM_INT, predicate=hb.HS_PRED_SQRT_LE_R,
                       flags=flags | hb.HS_FLAG_SORT_HITS)
        h.seed_projection(12345)
        h.load_fragments(codes)
        allp = h.bruteforce_codes(None, cap=1 << 22)
        st_all = h.stats()
        expl = h.bruteforce_codes(qcodes, cap=1 << 22)
        h.build_index()
        srch = h.search_codes(qcodes, cap=1 << 22)
        lab = h.cluster()
        res.append((hits_as_tuples(allp, False), hits_as_tuples(expl, False), hits_as_tuples(srch), lab,
                    st_all.n_candidates_tc, st_all.n_survivors))
        h.close()
    assert len(res[0][0]) > 0 and len(res[0][1]) > 0 and len(res[0][2]) > 0
    for r in res[1:]:
        assert r[0] == res[0][0] and r[1] == res[0][1] and r[2] == res[0][2] and np.array_equal(r[3], res[0][3])
    assert res[0][4] > 0 and res[2][4] == 0          # tensor filter in use / scalar only
    want = oracle.bruteforce_int(codes[:6000], None, R)
    sub = [t for t in res[0][0] if t[1] < 6000]       # (query, db_id, dist): all-pairs hits inside the first 6000
    assert sub == hits_as_tuples(want, False)


def test_search_int_metric(oracle):
    """LSH candidates verified with the integer window distance (V3)."""
    n, q, R = 20000, 200, 40
    codes = random_codes(n, 10, seed=71)
    qcodes = planted_queries(codes, q, seed=72)
    tab = oracle.coordinates(True)
    h, a, b = make(10, 4, 4, 50.0, float(R), metric=hb.HS_METRIC_BLOSUM_INT)
    h.load_fragments(codes)
    h.build_index()
    got = h.search_codes(qcodes)
    # oracle: candidate sets from the reference search with R = inf, then V3 on each candidate
    cand, _, _ = oracle.search(oracle.embed(codes, tab), oracle.embed(qcodes, tab), a, b, 50.0, 1e9)
    exp = []
    for c in cand:
        d = oracle.distance_int(qcodes[c["query"]], codes[c["db_id"]])
        if d <= R:
            exp.append((int(c["query"]), int(c["table_first"]), int(c["db_id"]), float(d)))
    assert len(exp) > 0
    assert hits_as_tuples(got) == exp
    h.close()


# ---------------------------------------------------------------- K5: cluster
@pytest.mark.parametrize("metric,R", [(hb.HS_METRIC_EUCLID_FP64, 25.0), (hb.HS_METRIC_BLOSUM_INT, 30.0)])
@pytest.mark.parametrize("K,W", [(4, 50.0), (8, 30.0)])
def test_cluster_partition(oracle, metric, R, K, W):
    n, L = 6000, 4
    codes = planted_families(n, 10, seed=81)
    tab = oracle.coordinates(True)
    h, a, b = make(10, K, L, W, R, metric=metric, predicate=hb.HS_PRED_SQRT_LE_R)
    h.load_fragments(codes)
    h.build_index()
    got = h.cluster()
    want, ne = oracle.cluster(codes, tab, a, b, W, R, metric=0 if metric == hb.HS_METRIC_EUCLID_FP64 else 1)
    assert ne > 0 and len(set(want.tolist())) < n
    assert np.array_equal(got, want)
    h.close()


# ---------------------------------------------------------------- E2: windows
def test_extract_windows(oracle):
    rng = np.random.default_rng(91)
    lens = rng.integers(0, 60, size=200)
    lens[3] = 9   # shorter than the window
    lens[4] = 10  # exactly one window
    start = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    residues = rng.integers(0, 20, size=int(start[-1]), dtype=np.uint8)
    for stride in (1, 3):
        h, a, b = make(10, 4, 4, 50.0, 30.0)
        nfrag, pos = h.extract_windows(residues, start, stride=stride)
        ocodes, opos = oracle.extract_windows(residues, start, 10, stride)
        assert nfrag == len(ocodes) and np.array_equal(pos, opos)
        got = h.hash(want_buckets=True)
        want = oracle.hash_codes(ocodes, oracle.coordinates(True), a, b, 50.0)
        assert np.array_equal(got, want)
        h.close()


# ---------------------------------------------------------------- tensor-core filter
@pytest.mark.parametrize("length,R,metric", [(10, 30.0, hb.HS_METRIC_EUCLID_FP64), (25, 60.0, hb.HS_METRIC_EUCLID_FP64),
                                             (32, 80.0, hb.HS_METRIC_EUCLID_FP64), (10, 40.0, hb.HS_METRIC_BLOSUM_INT),
                                             (3, 12.0, hb.HS_METRIC_EUCLID_FP64)])
def test_tensor_filter_equals_scalar_filter(oracle, length, R, metric):
    """The tcgen05 filter and the scalar filter must hand the exact stage survivor
    sets that produce identical hits (same order, same FP64 distances), and the
    tensor-core leg must actually have run."""
    n, q = 60000, 500
    codes = random_codes(n, length, seed=101)
    qcodes = planted_queries(codes, q, seed=102)
    res = []
    for flags in (hb.HS_FLAG_SORT_HITS, hb.HS_FLAG_SORT_HITS | hb.HS_FLAG_SCALAR_FILTER):
        h, a, b = make(length, 4, 4, 50.0, R, metric=metric, flags=flags)
        h.load_fragments(codes)
        h.build_index()
        got = h.search_codes(qcodes, cap=1 << 22)
        st = h.stats()
        if flags & hb.HS_FLAG_SCALAR_FILTER:
            assert st.n_candidates_tc == 0
        else:
            assert st.n_candidates_tc > 0
        res.append((hits_as_tuples(got), st.n_candidates))
        h.close()
    assert res[0][1] == res[1][1]
    assert len(res[0][0]) > 0 and res[0][0] == res[1][0]


def test_tensor_filter_full_query_tiles(oracle):
    """Buckets probed by more than 128 queries are split into several items; odd
    query counts exercise the N padding."""
    n, q, length = 40000, 1111, 10
    codes = random_codes(n, length, seed=111)
    qcodes = planted_queries(codes, q, seed=112)
    tab = oracle.coordinates(True)
    h, a, b = make(length, 2, 2, 80.0, 28.0)
    h.load_fragments(codes)
    h.build_index()
    got = h.search_codes(qcodes, cap=1 << 23)
    assert h.stats().n_candidates_tc > 0
    want, ts, _ = oracle.search(oracle.embed(codes, tab), oracle.embed(qcodes, tab), a, b, 80.0, 28.0, pred=0)
    assert len(want) > 0
    assert hits_as_tuples(got) == hits_as_tuples(want)
    h.close()


def test_bruteforce_euclid_all_pairs(oracle):
    """All pairs i<j of the DB under the Euclidean metric (the DB is its own query
    set: the tensor filter builds its query rows from the residue codes)."""
    n, length, R = 5000, 10, 22.0
    codes = planted_families(n, length, seed=121)
    tab = oracle.coordinates(True)
    h = hb.HSearch(length, 4, 4, 50.0, R, predicate=hb.HS_PRED_SQRT_LE_R)
    h.load_fragments(codes)
    got = h.bruteforce_codes(None, cap=1 << 22)
    assert h.stats().n_candidates_tc > 0
    pts = oracle.embed(codes, tab)
    want = oracle.bruteforce(pts, pts, R, pred=1, cap=1 << 24)
    want = want[want["query"] < want["db_id"]]
    assert len(want) > 0
    assert hits_as_tuples(got, False) == hits_as_tuples(want, False)
    h.close()


def test_tensor_filter_unrepresentable_queries(oracle):
    """Dense query points outside the FP16 range (or not finite) must not lose
    hits: such a query passes the whole bucket to the exact stage."""
    n, q, length, R = 30000, 64, 10, 30.0
    codes = random_codes(n, length, seed=131)
    qcodes = planted_queries(codes, q, seed=132)
    tab = oracle.coordinates(True)
    qpts = oracle.embed(qcodes, tab).copy()
    qpts[1, 3] = 1.0e6       # far away: no hits, but the row must not poison its tile
    qpts[2, 0] = np.inf
    qpts[3, 5] = np.nan
    qpts[4] *= 1.0 + 1e-9    # not a table row
    res = []
    for flags in (hb.HS_FLAG_SORT_HITS, hb.HS_FLAG_SORT_HITS | hb.HS_FLAG_SCALAR_FILTER):
        h, a, b = make(length, 2, 3, 80.0, R, flags=flags)
        h.load_fragments(codes)
        h.build_index()
        res.append(hits_as_tuples(h.search_points(qpts, cap=1 << 22)))
        h.close()
    assert len(res[0]) > 0 and res[0] == res[1]
    ok = np.ones(q, dtype=bool)
    ok[[1, 2, 3]] = False
    want, _, _ = oracle.search(oracle.embed(codes, tab), qpts[ok], a, b, 80.0, R, pred=0)
    remap = np.flatnonzero(ok)
    want_t = [(int(remap[qq]), t, d, x) for qq, t, d, x in hits_as_tuples(want)]
    assert [r for r in res[0] if ok[r[0]]] == want_t


def test_tensor_filter_huge_threshold(oracle):
    """R so large that every pair is a hit (the non-hit dump of
    motif_both_points_noLSH uses this): thresholds become infinite."""
    n, q, length = 3000, 40, 10
    codes = random_codes(n, length, seed=141)
    qcodes = random_codes(q, length, seed=142)
    tab = oracle.coordinates(True)
    h = hb.HSearch(length, 1, 1, 1.0, 1e300, predicate=hb.HS_PRED_SQRT_LE_R)
    h.load_fragments(codes)
    got = h.bruteforce_codes(qcodes, cap=n * q + 10)
    assert len(got) == n * q
    want = oracle.bruteforce(oracle.embed(codes, tab), oracle.embed(qcodes, tab), 1e300, pred=1, cap=n * q + 10)
    assert hits_as_tuples(got, False) == hits_as_tuples(want, False)
    h.close()


# ---------------------------------------------------------------- dense bucket ranks vs packed keys
@pytest.mark.parametrize("length,K,L,W,R", [(10, 4, 4, 50.0, 30.0), (10, 2, 6, 30.0, 25.0), (25, 3, 2, 80.0, 60.0),
                                            (10, 4, 20, 50.0, 30.0), (16, 4, 3, 50.0, 40.0)])
def test_rank_path_equals_packed_key_path(oracle, monkeypatch, length, K, L, W, R):
    """The u16 bucket-rank index (<= 65536 possible key strings per table) and the packed
    64-bit-key index must agree on keys, bucket order, table sizes and hits."""
    codes = planted_families(30011, length, seed=5)
    qcodes = planted_queries(codes, 200, seed=6)
    res = []
    for no_rank in ("0", "1"):
        monkeypatch.setenv("HS_NO_RANK", no_rank)
        h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS | hb.HS_FLAG_HASH_AUDIT)
        h.load_fragments(codes)
        h.hash()
        st = h.stats()
        assert st.residual_flips == 0
        if no_rank == "1":
            assert st.rank_path == 0
        elif length <= 10 or K <= 3:
            assert st.rank_path == 1   # few enough possible key strings: dense ranks in use
        keys = [h.keys(l) for l in range(L)]
        h.build_index()
        sizes = h.table_sizes()
        tabs = [h.table(l) for l in range(L)]
        hits = h.search_codes(qcodes)
        labels = h.cluster() if L <= 6 else None
        res.append((keys, sizes, tabs, hits_as_tuples(hits), labels))
        h.close()
    (k0, s0, t0, h0, l0), (k1, s1, t1, h1, l1) = res
    for l in range(L):
        assert np.array_equal(k0[l], k1[l])
        assert np.array_equal(t0[l][0], t1[l][0]) and np.array_equal(t0[l][1], t1[l][1])
    assert np.array_equal(s0, s1)
    assert len(h0) > 0 and h0 == h1
    if l0 is not None:
        assert np.array_equal(l0, l1)


@pytest.mark.parametrize("length,K,L,W,R", [(10, 16, 4, 200.0, 30.0), (10, 8, 3, 10.0, 30.0), (30, 8, 3, 50.0, 70.0),
                                            (10, 16, 5, 20.0, 30.0)])
def test_hashed_key_sort_equals_full_key_sort(oracle, monkeypatch, length, K, L, W, R):
    """Multi-word keys (K >= 8 / small W) are sorted on a 64-bit hash of the key string, one word
    instead of 2-4; every member's full key is checked against its bucket's.  The buckets (as sets of
    ascending id lists: the reference's HashTable is an unordered_map, lsh.hpp:51-59), the table sizes
    and the hits must equal the sort on all key words, also when a collision is reported (fallback)."""
    codes = planted_families(30011, length, seed=15)
    qcodes = planted_queries(codes, 300, seed=16)
    res = []
    for env in ({"HS_FORCE_HASH_SORT": "1"}, {"HS_NO_HASH_SORT": "1"}, {"HS_FORCE_HASH_COLLISION": "1"}, {}):
        for k in ("HS_NO_HASH_SORT", "HS_FORCE_HASH_COLLISION", "HS_FORCE_HASH_SORT"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
        h.load_fragments(codes)
        h.build_index()
        st = h.stats()
        assert st.key_words >= 2 and st.rank_path == 0
        assert st.hash_sort_fallbacks == (L if "HS_FORCE_HASH_COLLISION" in env else 0)
        sizes = h.table_sizes()
        buckets = []
        for l in range(L):
            ids, starts = h.table(l)
            assert starts[0] == 0 and starts[-1] == len(codes) and np.all(np.diff(starts.astype(np.int64)) > 0)
            bl = sorted(tuple(ids[starts[i]:starts[i + 1]].tolist()) for i in range(len(starts) - 1))
            assert all(list(t) == sorted(t) for t in bl)
            buckets.append(bl)
        hits = h.search_codes(qcodes)
        labels = h.cluster()
        res.append((sizes, buckets, hits_as_tuples(hits), labels, st.sort_passes))
        h.close()
    for r in res[1:]:
        assert np.array_equal(res[0][0], r[0])
        assert res[0][1] == r[1]
        assert len(r[2]) > 0 and res[0][2] == r[2]
        assert np.array_equal(res[0][3], r[3])
    assert res[0][4] <= 8 * L               # one word, <= 8 passes per table
    assert res[3][4] in (res[0][4], res[1][4])   # the default picks one of the two
    tab = oracle.coordinates(True)
    want, ts, _ = oracle.search(oracle.embed(codes, tab), oracle.embed(qcodes, tab), a, b, W, R, pred=0)
    assert np.array_equal(res[0][0], ts) and res[0][2] == hits_as_tuples(want)


def test_lazy_code_stores_and_filter_bypass(monkeypatch):
    """An index of tiny buckets (K = 16: millions of buckets of one or two members) is built without
    the bucket-ordered code stores; a search with few candidates sends them straight to the exact
    stage, a search with many (or hs_cluster) builds the stores first.  Hits and labels must equal
    the eager build's."""
    length, K, L, W, R = 10, 16, 3, 20.0, 30.0
    n = (1 << 20) + 12345
    codes = planted_families(200_000, length, seed=25)
    codes = np.concatenate([codes, random_codes(n - len(codes), length, seed=26)])
    qcodes = planted_queries(codes[:200_000], 2500, seed=27)
    res = []
    for env in ({"HS_NO_LAZY_STORES": "1"}, {}, {"HS_BYPASS_MAX": "10"}):
        for k in ("HS_NO_LAZY_STORES", "HS_BYPASS_MAX"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
        h.load_fragments(codes)
        h.build_index()
        permute_ms = h.stats().ms_permute
        hits = h.search_codes(qcodes)
        st = h.stats()
        hits2 = h.search_codes(qcodes[:100])
        labels = h.cluster()
        res.append((hits_as_tuples(hits), hits_as_tuples(hits2), labels, permute_ms, st.n_candidates, st.n_survivors,
                    st.n_work_items))
        h.close()
    assert len(res[0][0]) > 0
    for r in res[1:]:
        assert r[0] == res[0][0] and r[1] == res[0][1] and np.array_equal(r[2], res[0][2])
        assert r[3] < res[0][3]                    # no store build inside hs_build_index (only the fragment records)
    assert res[1][4] == res[1][5] and res[1][6] == 0   # bypass: every candidate is a survivor, no filter work list
    assert res[2][5] <= res[2][4] and res[2][6] > 0    # stores built on demand, filter ran
    assert res[0][4] == res[1][4] == res[2][4]


def test_search_pipelined_blocks_equal_single_pass(oracle, monkeypatch):
    """>= 2048 queries with a host hit buffer: the survivors of the (single) filter pass are split
    by query block, and each block's sorted hits are copied out while the next block is verified
    and sorted: same hits, same order as the unpipelined path, also at the capacity limit."""
    length, K, L, W, R = 10, 4, 4, 50.0, 30.0
    codes = random_codes(40000, length, seed=21)
    qcodes = planted_queries(codes, 2600, seed=22, frac=0.5)
    monkeypatch.setenv("HS_NO_PIPELINE", "1")   # read once, by hs_create
    h1, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
    h1.load_fragments(codes)
    h1.build_index()
    single = h1.search_codes(qcodes)
    h1.close()
    monkeypatch.setenv("HS_NO_PIPELINE", "0")
    h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
    h.load_fragments(codes)
    h.build_index()
    blocks = h.search_codes(qcodes)
    assert len(single) > 1000
    assert np.array_equal(single, blocks)
    tab = oracle.coordinates(True)
    want, _, _ = oracle.search(oracle.embed(codes, tab), oracle.embed(qcodes[:300], tab), a, b, W, R)
    got300 = blocks[blocks["query"] < 300]
    assert hits_as_tuples(got300) == hits_as_tuples(want)
    # capacity protocol: too small a buffer is reported, and growing it gives the same hits
    import ctypes
    with pytest.raises(hb.HsError):
        h._call_hits(h.lib.hs_search_codes, np.ascontiguousarray(qcodes), ctypes.c_uint8, len(qcodes),
                     len(single) // 2, grow=False)
    assert np.array_equal(h.search_codes(qcodes, cap=len(single) // 3), single)
    h.close()


def test_load_fragments_overlapped_hash(oracle, monkeypatch):
    """With a projection set, hs_load_fragments copies the DB in blocks and hashes each block as
    it arrives: same ranks / keys / index / hits as the plain load followed by hs_hash."""
    length, K, L, W, R = 10, 4, 4, 50.0, 30.0
    n = 8 * 768 * 256 + 12345
    codes = random_codes(n, length, seed=31)
    qcodes = planted_queries(codes[:50000], 64, seed=32)
    res = []
    for no_overlap in ("1", "0"):
        monkeypatch.setenv("HS_NO_LOAD_OVERLAP", no_overlap)
        h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
        h.load_fragments(codes)
        h.build_index()
        st = h.stats()
        keys = h.keys(2)
        ids, starts = h.table(1)
        hits = h.search_codes(qcodes)
        res.append((keys, ids, starts, hits, st.guard_hits))
        h.close()
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    assert len(res[0][3]) > 0 and np.array_equal(res[0][3], res[1][3])
    assert res[0][4] == res[1][4]
    want = oracle.hash_codes(codes[:2000], oracle.coordinates(True), a, b, W)
    strings = oracle.key_strings(want)
    expect = np.stack([hb.pack_key_string(s, 1) for s in strings[:, 2]])
    assert np.array_equal(res[1][0][:2000], expect)


def test_cluster_sharded_path_on_one_rank(oracle):
    """The sharded cluster path (per-table in-bucket pair search on the bucket owner, edge
    union through hs_union_find) with world size 1 must give the labels of hs_cluster / the oracle."""
    from hsearch_b200 import dist as hdist
    length, K, L, W, R = 10, 4, 3, 50.0, 25.0
    codes = planted_families(5003, length, seed=41)
    h, a, b = make(length, K, L, W, R, predicate=hb.HS_PRED_SQRT_LE_R)
    h.load_fragments(codes)
    h.build_index()
    direct = h.cluster()
    key_fn, edges_fn, union_fn = hdist.gpu_cluster_callbacks(h, a, b, codes)
    got = hdist.cluster_sharded(codes, 0, len(codes), L, key_fn, edges_fn, union_fn)
    want, ne = oracle.cluster(codes, oracle.coordinates(True), a, b, W, R)
    assert ne > 0 and np.array_equal(direct, want)
    assert np.array_equal(got, want)
    # hs_union_find alone against the oracle's UnionFind (SURVEY 8c known answer)
    eu, ev = np.array([0, 20, 10, 50, 70]), np.array([10, 30, 30, 60, 50])
    lab = h.union_find(80, eu, ev)
    assert lab[[0, 10, 20, 30]].tolist() == [0, 0, 0, 0] and lab[40] == 40 and lab[[50, 60, 70]].tolist() == [50] * 3
    with pytest.raises(hb.HsError):
        h.union_find(10, np.array([3]), np.array([11]))
    h.close()


@pytest.mark.parametrize("W", [50.0, 20.0])
def test_blocked_gather_equals_per_table_gather(oracle, monkeypatch, W):
    """Above 128 MB of fragment records the bucket-order code stores of all tables are built by
    one L2-blocked gather; the searches that stream those stores must return the same hits
    as with the per-table gather (rank path at W = 50, packed-key path at W = 20)."""
    length, K, L, R = 10, 4, 4, 30.0
    n = 4_300_003
    codes = random_codes(n, length, seed=51)
    qcodes = planted_queries(codes[:100000], 300, seed=52, frac=0.5)
    res = []
    for off in ("1", "0"):
        monkeypatch.setenv("HS_NO_BLOCKED_GATHER", off)
        h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
        h.load_fragments(codes)
        h.build_index()
        res.append((h.search_codes(qcodes), h.table_sizes(), h.stats().rank_path))
        h.close()
    assert res[0][2] == res[1][2] == (1 if W == 50.0 else 0)
    assert np.array_equal(res[0][1], res[1][1])
    assert len(res[0][0]) > 300 and np.array_equal(res[0][0], res[1][0])
    # and against the oracle on a slice of the queries (brute force over the hits' distances)
    tab = oracle.coordinates(True)
    hits = res[1][0]
    sub = hits[hits["query"] < 20]
    d = oracle.embed(codes[sub["db_id"]], tab) - oracle.embed(qcodes[sub["query"]], tab)
    assert np.all((d * d).sum(axis=1) <= R * R * (1 + 1e-12))


def _greedy_reference(codes, table, a, b, W, R, oracle):
    """Clustering() of hclust2.cpp:86-151 restated in Python on top of the oracle's hash and
    distance (small inputs only)."""
    n = len(codes)
    L = a.shape[0]
    pts = oracle.embed(codes, table)
    strings = oracle.key_strings(oracle.hash_codes(codes, table, a, b, W))
    merged = np.zeros(n, dtype=np.uint8)
    center = np.arange(n, dtype=np.uint32)
    rnd = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    for l in range(L):
        buckets = {}
        for i in range(n):
            if merged[i] != 2:
                buckets.setdefault(strings[i, l], []).append(i)
        for ids in buckets.values():
            centers = [i for i in ids if merged[i] == 1]
            for i in ids:
                if merged[i] == 0:
                    for c in centers:
                        if not (np.sqrt(oracle.dist2(pts[i], pts[c])) > R):
                            merged[c] = 1
                            merged[i] = 2
                            center[i] = c
                            rnd[i] = l
                            break
                if merged[i] == 0:
                    centers.append(i)
    return center, rnd, merged


@pytest.mark.parametrize("length,K,L,W,R,n", [(10, 4, 4, 50.0, 25.0, 1500), (10, 2, 3, 20.0, 30.0, 1200),
                                              (25, 3, 2, 80.0, 60.0, 800), (8, 4, 1, 50.0, 1e9, 300)])
def test_greedy_cluster_matches_restatement(oracle, length, K, L, W, R, n):
    """hs_greedy_cluster against a direct restatement of hclust2's Clustering() (the CLI test
    compares with the reference binary itself); R = 1e9 joins every bucket to its first member."""
    codes = planted_families(n, length, seed=61)
    h, a, b = make(length, K, L, W, R, table_variant=hb.HS_TABLE_FULL, predicate=hb.HS_PRED_SQRT_LE_R, flags=0)
    h.load_fragments(codes)
    h.build_index()
    center, rnd, merged = h.greedy_cluster()
    wc, wr, wm = _greedy_reference(codes, oracle.coordinates(False), a, b, W, R, oracle)
    assert np.array_equal(merged, wm)
    assert np.array_equal(center, wc)
    assert np.array_equal(rnd, wr)
    assert (merged == 2).sum() > 0
    h.close()


def test_search_pipelined_dense_queries_and_small_capacity(oracle, monkeypatch):
    """The query-block pipeline with dense (non-residue) centres and a hit buffer that overflows
    in a middle block: the needed size is reported and the retry returns the single-pass hits."""
    length, K, L, W, R = 10, 4, 4, 50.0, 32.0
    codes = random_codes(30000, length, seed=71)
    tab = oracle.coordinates(True)
    rng = np.random.default_rng(72)
    q = oracle.embed(planted_queries(codes, 2300, seed=73, frac=0.8), tab)
    q[::3] += rng.normal(0.0, 0.5, size=q[::3].shape)      # every third centre is not a residue string
    monkeypatch.setenv("HS_NO_PIPELINE", "1")   # read once, by hs_create
    h1, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
    h1.load_fragments(codes)
    h1.build_index()
    single = h1.search_points(q)
    h1.close()
    monkeypatch.setenv("HS_NO_PIPELINE", "0")
    h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
    h.load_fragments(codes)
    h.build_index()
    assert len(single) > 500
    assert np.array_equal(h.search_points(q), single)
    assert np.array_equal(h.search_points(q, cap=len(single) // 2), single)
    want, _, _ = oracle.search(oracle.embed(codes, tab), q[:200], a, b, W, R)
    assert hits_as_tuples(single[single["query"] < 200]) == hits_as_tuples(want)
    h.close()


def test_bruteforce_points_dev_matches_host_variant(oracle):
    """Device-resident brute force (queries and hits on the device; cap == 0 gives the count only)."""
    torch = pytest.importorskip("torch")
    length, R = 10, 30.0
    codes = random_codes(20000, length, seed=81)
    tab = oracle.coordinates(True)
    q = oracle.embed(planted_queries(codes, 50, seed=82, frac=0.6), tab)
    h, a, b = make(length, 4, 4, 50.0, R, flags=hb.HS_FLAG_SORT_HITS)
    h.load_fragments(codes)
    want = h.bruteforce_points(q)
    stream = torch.cuda.ExternalStream(h.stream_ptr())
    with torch.cuda.stream(stream):
        qd = torch.from_numpy(q).cuda()
        torch.cuda.synchronize()
        n = h.bruteforce_points_dev(qd.data_ptr(), len(q), 0, 0)
        assert n == len(want) > 0
        buf = torch.zeros(n * 24, dtype=torch.uint8, device="cuda")
        assert h.bruteforce_points_dev(qd.data_ptr(), len(q), buf.data_ptr(), n) == n
        torch.cuda.synchronize()
    got = np.frombuffer(buf.cpu().numpy().tobytes(), dtype=hb.HIT_DTYPE)
    assert np.array_equal(got, want)
    ref = oracle.bruteforce(oracle.embed(codes, tab), q, R, pred=0)
    assert hits_as_tuples(got, False) == hits_as_tuples(ref, False)
    h.close()


def test_blocked_gather_long_fragments(oracle, monkeypatch):
    """len 25: two 16-byte words of codes per record (48-byte records) through the blocked gather,
    the tensor filter and the exact stage."""
    length, K, L, W, R = 25, 4, 3, 120.0, 62.0
    n = 1_500_007
    codes = random_codes(n, length, seed=91)
    qcodes = planted_queries(codes[:50000], 200, seed=92, frac=0.7)
    res = []
    for off in ("1", "0"):
        monkeypatch.setenv("HS_NO_BLOCKED_GATHER", off)
        h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
        h.load_fragments(codes)
        h.build_index()
        res.append((h.search_codes(qcodes), h.table_sizes()))
        h.close()
    assert np.array_equal(res[0][1], res[1][1])
    assert len(res[0][0]) > 100 and np.array_equal(res[0][0], res[1][0])
    tab = oracle.coordinates(True)
    sub = res[1][0][res[1][0]["query"] < 40]
    want, _, _ = oracle.search(oracle.embed(codes[:200000], tab), oracle.embed(qcodes[:40], tab), a, b, W, R)
    got_sub = sub[sub["db_id"] < 200000]
    assert hits_as_tuples(got_sub) == hits_as_tuples(want)


def test_cluster_large_buckets_in_query_chunks(oracle, monkeypatch):
    """A large bucket is self-joined in chunks of its members (bounded survivor buffer, what lets the
    50 M configuration run): 1 k-member chunks over ~9 k-member buckets give the same edges and labels as
    one pass per bucket."""
    length, K, L, W, R = 10, 4, 3, 50.0, 25.0
    codes = planted_families(115000, length, seed=96)
    res = []
    for chunk in ("1024", "1000000"):
        monkeypatch.setenv("HS_SELFJOIN_CHUNK", chunk)     # read by hs_create
        h, a, b = make(length, K, L, W, R, predicate=hb.HS_PRED_SQRT_LE_R)
        h.load_fragments(codes)
        h.build_index()
        res.append((h.cluster(), h.stats().as_dict()))
        h.close()
    assert res[0][1]["n_candidates_tc"] == res[1][1]["n_candidates_tc"] > 1e7
    assert res[0][1]["n_candidates"] == res[1][1]["n_candidates"]
    assert res[0][1]["n_edges"] == res[1][1]["n_edges"] > 0
    assert np.array_equal(res[0][0], res[1][0])


@pytest.mark.parametrize("R,pred", [(25.0, hb.HS_PRED_SQRT_LE_R), (30.0, hb.HS_PRED_D2_LE_R2)])
def test_cluster_large_buckets_through_tensor_filter(oracle, R, pred):
    """Buckets of >= 8192 members are self-joined by the tcgen05 filter (queries = the bucket's own
    members): same partition as the oracle and as the scalar-filter path."""
    length, K, L, W = 10, 4, 3, 50.0
    codes = planted_families(115000, length, seed=95)    # W = 50: the largest buckets hold ~8 % of the DB
    res = []
    for flags in (0, hb.HS_FLAG_SCALAR_FILTER):
        h, a, b = make(length, K, L, W, R, predicate=pred, flags=flags)
        h.load_fragments(codes)
        h.build_index()
        res.append((h.cluster(), h.stats().as_dict()))
        h.close()
    assert res[0][1]["n_candidates_tc"] > 1e7 and res[1][1]["n_candidates_tc"] == 0
    assert res[0][1]["n_edges"] == res[1][1]["n_edges"] > 0
    assert np.array_equal(res[0][0], res[1][0])
    want, ne = oracle.cluster(codes, oracle.coordinates(True), a, b, W, R, metric=0) if pred == hb.HS_PRED_SQRT_LE_R \
        else (None, 0)
    if want is not None:
        assert np.array_equal(res[0][0], want)


# ---------------------------------------------------------------- R1: recall evaluation on the device
def test_evaluate_recall_golden(oracle):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "evaluate_golden.npz"))
    R, Q = float(g["R"]), int(g["Q"])
    h = hb.HSearch(10, 4, 4, 50.0, R)
    r = h.evaluate_recall(g["truth"], g["found"], Q)
    # counts and bins exact; the weighted sums are added in a parallel order (tolerance 1e-12)
    assert np.array_equal(r["tp_bin"], g["tp_bin"]) and np.array_equal(r["fn_bin"], g["fn_bin"])
    assert r["n_tp"] == int(g["tp_bin"].sum()) and r["n_fn"] == int(g["fn_bin"].sum()) and r["n_extra"] == 30
    e = oracle.evaluate(g["truth"], g["found"], R)
    assert abs(r["tp"] - e["tp"]) <= 1e-12 * e["tp"] and abs(r["fn"] - e["fn"]) <= 1e-12 * e["fn"]
    assert abs(r["recall"] - float(g["recall"])) <= 1e-12
    # truth in any order gives the same answer; found out of order is refused
    perm = np.random.default_rng(3).permutation(len(g["truth"]))
    r2 = h.evaluate_recall(g["truth"][perm], g["found"], Q)
    assert r2["n_tp"] == r["n_tp"] and np.array_equal(r2["tp_bin"], r["tp_bin"])
    with pytest.raises(hb.HsError):
        h.evaluate_recall(g["truth"], g["found"][::-1], Q)
    bad = g["truth"].copy()
    bad["dist2"][7] = (R + 0.2) ** 2
    with pytest.raises(hb.HsError):
        h.evaluate_recall(bad, g["found"], Q)
    # empty lists
    z = h.evaluate_recall(g["truth"][:0], g["found"], Q)
    assert z["n_tp"] == 0 and z["n_fn"] == 0 and z["n_extra"] == len(g["found"])
    z = h.evaluate_recall(g["truth"], g["found"][:0], Q)
    assert z["n_tp"] == 0 and z["n_fn"] == len(g["truth"]) and z["recall"] == 0.0
    h.close()


@pytest.mark.parametrize("length,W,R", [(10, 50.0, 30.0), (10, 20.0, 30.0), (25, 50.0, 55.0)])
def test_evaluate_recall_of_a_search(oracle, length, W, R):
    # the reference's pipeline: brute force = ground truth, LSH search = found, evaulate() joins them
    n, q = 30000, 200
    codes = random_codes(n, length, seed=81)
    qcodes = planted_queries(codes, q, seed=82, frac=0.8, max_sub=4)
    h, a, b = make(length, 4, 4, W, R, flags=hb.HS_FLAG_SORT_HITS)
    h.load_fragments(codes)
    h.build_index()
    found = h.search_codes(qcodes)
    truth = h.bruteforce_codes(qcodes)
    assert 0 < len(found) <= len(truth)
    r = h.evaluate_recall(truth, found, q)
    e = oracle.evaluate(truth, found, R)
    assert (r["n_tp"], r["n_fn"], r["n_extra"]) == (e["n_tp"], e["n_fn"], e["n_extra"]) == (len(found), len(truth) - len(found), 0)
    assert np.array_equal(r["tp_bin"], e["tp_bin"]) and np.array_equal(r["fn_bin"], e["fn_bin"])
    assert abs(r["tp"] - e["tp"]) <= 1e-12 * max(e["tp"], 1.0) and abs(r["fn"] - e["fn"]) <= 1e-12 * max(e["fn"], 1.0)
    # device-resident variant
    import torch
    td = torch.from_numpy(truth.view(np.uint8).copy()).cuda()
    fd = torch.from_numpy(found.view(np.uint8).copy()).cuda()
    torch.cuda.synchronize()
    rd = h.evaluate_recall_dev(td.data_ptr(), len(truth), fd.data_ptr(), len(found), q)
    assert rd["tp"] == r["tp"] and rd["fn"] == r["fn"] and np.array_equal(rd["tp_bin"], r["tp_bin"])
    h.close()
