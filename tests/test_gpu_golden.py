"""GPU tests: the CUDA path against the committed golden fixtures, which were
produced by the reference's own code (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import hsearch_b200 as hb
from tests.util import hits_as_tuples

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_hash_against_reference_golden(gold):
    codes = gold["hash_codes"]
    for tag, variant in (("p6", hb.HS_TABLE_PRINT6), ("full", hb.HS_TABLE_FULL)):
        for (K, L, W) in [(4, 4, 50.0), (4, 4, 4.0), (16, 2, 10.0)]:
            h = hb.HSearch(10, K, L, W, 30.0, table_variant=variant)
            h.seed_projection(777)
            h.load_fragments(codes)
            got = h.hash(want_buckets=True)
            key = f"hash_{tag}_K{K}_L{L}_W{W:g}"
            assert np.array_equal(got, gold[key + "_buckets"])
            kw = h.stats().key_words
            for l in range(L):
                expect = np.stack([hb.pack_key_string(s, kw) for s in gold[key + "_keys"][:, l].tolist()])
                assert np.array_equal(h.keys(l), expect)
            h.close()


def test_search_against_reference_golden(gold):
    for W in (20.0, 50.0):
        h = hb.HSearch(10, 4, 4, W, 30.0)
        h.seed_projection(12345)
        h.load_fragments(gold["search_db"])
        h.build_index()
        got = h.search_codes(gold["search_q"])
        ref = gold[f"search_W{W:g}_hits"]
        assert hits_as_tuples(got, False) == hits_as_tuples(ref, False)
        assert np.array_equal(h.table_sizes(), gold[f"search_W{W:g}_tsizes"])
        printed = np.array([float("%g" % v) for v in np.sqrt(got["dist2"])])
        assert np.array_equal(printed, gold[f"search_W{W:g}_printed"])
        h.close()


def test_bruteforce_against_reference_golden(gold):
    h = hb.HSearch(10, 4, 4, 50.0, 30.0, predicate=hb.HS_PRED_SQRT_LE_R)
    h.load_fragments(gold["search_db"])
    got = h.bruteforce_codes(gold["search_q"])
    ref = gold["brute_hits"]
    assert np.array_equal(got["query"], ref["query"]) and np.array_equal(got["db_id"], ref["db_id"])
    assert np.array_equal(np.sqrt(got["dist2"]), ref["dist2"])
    h.close()
