"""CPU tests: the oracle restatement (oracle/hs_oracle.c) against the golden
vectors produced by the reference itself (tests/golden/make_golden.py), the
known-answer vectors of SURVEY.md 8c, and the reference's static tables."""
import os

import numpy as np
import pytest

from tests.util import hits_as_tuples

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_known_answer_vectors_from_survey(oracle):
    # LSH(dim=80, K=4, W=50.0), engine seed 12345
    a, b = oracle.lsh_generate(12345, 80, 4, 50.0)
    assert a[0, 0] == 0.11176354368256498 and a[0, 1] == -0.59065862504305633
    assert a[0, 2] == -0.62564101996406651 and a[0, 3] == 0.58237074530768129
    assert a[0, 79] == 0.86829835845012571 and b[0] == 9.0812367446508979
    assert a[1, 0] == 0.65709572028881058 and b[3] == 48.866218389250967
    base = oracle.base()
    codes = np.array([[base[ord(c) - 65] for c in "ARNDCQEGHI"]], dtype=np.uint8)
    pt = oracle.embed(codes, oracle.coordinates())[0]
    dots = [float(np.float64(0) + sum_seq(pt, a[k])) for k in range(4)]
    assert dots == [15.669676326315257, 51.285231958298048, 15.847428872456565, 27.189868212478299]
    A, B = oracle.lsh_tables(12345, 80, 4, 1, 50.0)
    bk = oracle.hash_codes(codes, oracle.coordinates(), A, B, 50.0)
    assert bk.reshape(-1).tolist() == [0, 1, 1, 1]
    assert oracle.key_strings(bk)[0, 0] == "0111"
    # UnionFind example: partition {0,10,20,30},{40},{50,60,70}
    lab = oracle.union_find_labels(8, [0, 2, 1, 5, 7], [1, 3, 3, 6, 5])
    assert lab.tolist() == [0, 0, 0, 0, 4, 5, 5, 5]
    # KLSH(512,16,0.2)
    w, t, bb = oracle.klsh_generate()
    p = np.zeros(512)
    p[3], p[77], p[500] = 2, 1, 5
    assert oracle.klsh_hash(p, w, t, bb) == 17156
    assert oracle.klsh_hash(np.ones(512), w, t, bb) == 21252


def sum_seq(p, a):
    s = 0.0
    for x, y in zip(p.tolist(), a.tolist()):
        s += x * y
    return s


def test_static_tables(oracle):
    d = oracle.blosum_metric()
    assert np.array_equal(d, d.T) and d.max() == 26 and np.all(np.diag(d) == 0)
    assert oracle.triangle_violations(d) == 0
    # first rows of D as printed in IGC/distance2coordinate/BLOSUM.m:3-4
    assert d[0].tolist() == [0, 11, 14, 14, 13, 11, 11, 10, 16, 10, 10, 11, 11, 14, 13, 6, 9, 21, 15, 8]
    assert d[1].tolist() == [11, 0, 11, 15, 20, 8, 10, 15, 13, 15, 13, 6, 12, 17, 16, 11, 12, 22, 16, 15]
    # coordinates: squared distances reproduce DISTANCE_SQUARE (util.hpp:43-64) samples
    c = oracle.coordinates()
    dsq = ((c[:, None, :] - c[None, :, :]) ** 2).sum(-1)
    assert abs(dsq[0, 1] - 131.470960) < 1e-5 and abs(dsq[17, 3] - 676.000004) < 1e-5
    assert abs(dsq[9, 19] - 8.786247) < 1e-5
    # print6 table = 6 significant digits
    c6 = oracle.coordinates(print6=True)
    assert c6[17, 0] == 13.5924 and c[17, 0] == 13.592409
    assert np.max(np.abs(c6 - c) / np.abs(c)) < 5e-6 and np.sum(c6 != c) > 100
    # E<->Q swap of ProteinDB (protein.hpp:58-64)
    assert oracle.proteindb_code("E") == oracle.base()[ord("Q") - 65]
    assert oracle.proteindb_code("Q") == oracle.base()[ord("E") - 65]
    assert oracle.proteindb_code("A") == 0 and oracle.proteindb_code("B") == -1


def test_projection_golden(oracle, gold):
    for name in ("proj_a", "proj_b", "proj_c"):
        seed, dim, K, W = gold[name + "_args"]
        a, b = oracle.lsh_generate(int(seed), int(dim), int(K), float(W))
        assert np.array_equal(a, gold[name + "_a"]) and np.array_equal(b, gold[name + "_b"])


def test_hash_golden(oracle, gold):
    codes = gold["hash_codes"]
    for tag, print6 in (("p6", True), ("full", False)):
        tab = oracle.coordinates(print6)
        for (K, L, W) in [(4, 4, 50.0), (4, 4, 4.0), (16, 2, 10.0)]:
            a, b = oracle.lsh_tables(777, 80, K, L, W)
            got = oracle.hash_codes(codes, tab, a, b, W)
            key = f"hash_{tag}_K{K}_L{L}_W{W:g}"
            assert np.array_equal(got, gold[key + "_buckets"])
            assert np.array_equal(oracle.key_strings(got).astype("U200"), gold[key + "_keys"])


def test_search_golden(oracle, gold):
    tab = oracle.coordinates(True)
    db, q = oracle.embed(gold["search_db"], tab), oracle.embed(gold["search_q"], tab)
    for W in (20.0, 50.0):
        a, b = oracle.lsh_tables(12345, 80, 4, 4, W)
        hits, ts, _ = oracle.search(db, q, a, b, W, 30.0, pred=0)
        ref = gold[f"search_W{W:g}_hits"]
        assert len(ref) > 0
        assert hits_as_tuples(hits, False) == hits_as_tuples(ref, False)
        assert np.array_equal(ts, gold[f"search_W{W:g}_tsizes"])
        # the text the reference printed is sqrt(d2) at 6 significant digits
        printed = gold[f"search_W{W:g}_printed"]
        assert np.array_equal(np.array([float("%g" % v) for v in np.sqrt(hits["dist2"])]), printed)


def test_bruteforce_golden(oracle, gold):
    tab = oracle.coordinates(True)
    db, q = oracle.embed(gold["search_db"], tab), oracle.embed(gold["search_q"], tab)
    hits = oracle.bruteforce(db, q, 30.0, pred=1)
    ref = gold["brute_hits"]
    assert len(ref) > 0
    assert np.array_equal(hits["query"], ref["query"]) and np.array_equal(hits["db_id"], ref["db_id"])
    assert np.array_equal(np.sqrt(hits["dist2"]), ref["dist2"])  # reference harness stores PairwiseDistance (sqrt)


def test_union_find_golden(oracle, gold):
    n = int(gold["uf_n"][0])
    lab = oracle.union_find_labels(n, gold["uf_eu"], gold["uf_ev"])
    roots = gold["uf_roots"]
    # same partition (root identity is edge-order dependent, the partition is not)
    canon = {}
    for i, r in enumerate(roots.tolist()):
        canon.setdefault(r, i)
    assert lab.tolist() == [canon[r] for r in roots.tolist()]


def test_klsh_golden(oracle, gold):
    w, t, b = oracle.klsh_generate()
    for s, hv, ki in zip(gold["klsh_seqs"].tolist(), gold["klsh_hash"].tolist(), gold["kmer2int"].tolist()):
        f = oracle.kmer3_features(s)
        assert oracle.klsh_hash(f, w, t, b) == hv
        assert int(np.argmax(oracle.kmer3_features(s[:3]))) == ki


def test_weight_golden(oracle, gold):
    for d, w in zip(gold["weight_d"].tolist(), gold["weight_w"].tolist()):
        assert oracle.weight(d, 30.0) == w


def test_orf6(oracle):
    # orf.cc:39-74: each frame cut at the first stop, kept if >= 6 aa
    dna = "ATGGCCATTGTAATGGGCCGCTGAAAGGGTGCCCGATAG"
    frames = oracle.orf6(dna)
    assert frames[0] == "MAIVMGR"  # frame 0 stops at TGA
    assert all(len(f) >= 6 and "*" not in f for f in frames)


def test_windows_and_protein_id(oracle):
    start = np.array([0, 5, 17, 17, 40], dtype=np.uint32)
    residues = (np.arange(40) % 20).astype(np.uint8)
    codes, pos = oracle.extract_windows(residues, start, 10, 1)
    assert pos.tolist() == [5, 6, 7] + list(range(17, 31))
    assert np.array_equal(codes[0], residues[5:15])
    assert [oracle.protein_id(start, p) for p in (0, 4, 5, 16, 17, 39)] == [0, 0, 1, 1, 3, 3]


def test_reference_agrees_with_restatement_live(oracle, reference):
    """Where oracle/_ref is present: a fresh (non-golden) comparison."""
    from tests.util import planted_queries, random_codes
    db = random_codes(4000, 10, seed=7)
    qc = planted_queries(db, 60, seed=8)
    tab = oracle.coordinates(True)
    a, b = oracle.lsh_tables(4242, 80, 4, 4, 50.0)
    h, ts, _ = oracle.search(oracle.embed(db, tab), oracle.embed(qc, tab), a, b, 50.0, 30.0)
    rh, _, rts, _ = reference.search(oracle.embed(db, tab), oracle.embed(qc, tab), 4, 4, 50.0, 30.0, 4242)
    assert hits_as_tuples(h, False) == hits_as_tuples(rh, False) and np.array_equal(ts, rts)


# ---------------------------------------------------------------- R1: recall evaluation
EVAL_GOLD = os.path.join(os.path.dirname(__file__), "golden", "evaluate_golden.npz")


def test_evaluate_golden(oracle):
    # evaulate() of the reference (motif_both_points.cpp:100-165) on seeded lists: same sequential
    # sums, so the recall is bit-identical; bins are the rows of <out>.accuracy.txt
    g = np.load(EVAL_GOLD)
    e = oracle.evaluate(g["truth"], g["found"], float(g["R"]))
    assert e["recall"] == float(g["recall"])
    assert np.array_equal(e["tp_bin"], g["tp_bin"]) and np.array_equal(e["fn_bin"], g["fn_bin"])
    assert e["n_tp"] == int(g["tp_bin"].sum()) and e["n_fn"] == int(g["fn_bin"].sum())
    assert e["n_extra"] == len(g["found"]) - e["n_tp"] == 30


def test_evaluate_against_reference_itself(oracle, reference):
    from tests.golden.make_golden import evaluate_lists
    for seed, n in [(11, 1), (12, 257), (13, 3000)]:
        t, f, Q, R = evaluate_lists(seed=seed, n=n, Q=20, N=500, R=26.5)
        rv, rows = reference.evaluate(t, f, R)
        e = oracle.evaluate(t, f, R)
        assert e["recall"] == rv
        assert sum(int(r[0]) < 500 for r in rows) == int(((e["tp_bin"] + e["fn_bin"]) > 0).sum())


def test_evaluate_rejects_distance_above_threshold(oracle):
    g = np.load(EVAL_GOLD)
    t = g["truth"].copy()
    t["dist2"][5] = (float(g["R"]) + 0.2) ** 2  # weight() prints "err" and exits (:67-70)
    with pytest.raises(ValueError):
        oracle.evaluate(t, g["found"], float(g["R"]))
