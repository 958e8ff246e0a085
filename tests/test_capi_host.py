"""CPU tests of the C-ABI library: it loads, exports every symbol the header
declares, its host-side functions agree with the oracle, and compute entry
points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import hsearch_b200 as hb
from hsearch_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    header = open(os.path.join(ROOT, "include", "hsearch_b200.h")).read()
    declared = set(re.findall(r"\b(hs_[a-z0-9_]+)\s*\(", header))
    assert declared, "header parse failed"
    assert declared == set(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layouts_match_header():
    assert C.sizeof(capi.Params) == 48
    assert capi.HIT_DTYPE.itemsize == 24
    assert C.sizeof(capi.Stats) == 9 * 8 + 4 * 4 + 15 * 4 + 4 + 8 + 8 + 8 + 8  # 15 floats, padded to 8, four more u64


def test_host_tables_match_oracle(oracle):
    assert np.array_equal(hb.coordinates(hb.HS_TABLE_FULL), oracle.coordinates())
    assert np.array_equal(hb.coordinates(hb.HS_TABLE_PRINT6), oracle.coordinates(print6=True))
    assert np.array_equal(hb.blosum_metric(), oracle.blosum_metric())
    lib = capi.load()
    base = oracle.base()
    for i in range(26):
        ch = chr(65 + i)
        assert lib.hs_letter_to_code(ch.encode()) == base[i]
        assert lib.hs_proteindb_code(ch.encode()) == oracle.proteindb_code(ch)
    assert hb.encode(["ARNDCQEGHILKMFPSTWYV"]).tolist() == [list(range(20))]
    assert hb.AA_ORDER == "".join(chr(65 + int(np.where(base == c)[0][0])) for c in range(20))


def test_projection_generator_matches_oracle(oracle):
    for seed, dim, K, L, W in [(12345, 80, 4, 4, 50.0), (7, 200, 16, 3, 4.0), (2 ** 31 - 1, 8, 1, 2, 1.0)]:
        a, b = hb.generate_projection(seed, dim, K, L, W)
        oa, ob = oracle.lsh_tables(seed, dim, K, L, W)
        assert np.array_equal(a, oa) and np.array_equal(b, ob)


def test_pack_key_string():
    assert hb.pack_key_string("0", 1)[0] == 1
    assert hb.pack_key_string("0111", 1)[0] == 0x1222
    assert hb.pack_key_string("-12", 1)[0] == 0xB23
    # (1,23) and (12,3) concatenate to the same string -> same key (lsh.hpp:51-59)
    assert hb.pack_key_string("1" + "23", 1)[0] == hb.pack_key_string("12" + "3", 1)[0]
    w = hb.pack_key_string("1234567890-1234567", 2)
    assert w[1] == 0x23 and w[0] == 0x456789A1B2345678
    with pytest.raises(hb.HsError):
        hb.pack_key_string("1" * 17, 1)


@pytest.mark.skipif(capi.load().hs_device_available() == 1, reason="a B200 is present")
def test_no_cpu_fallback():
    with pytest.raises(hb.HsError) as e:
        hb.HSearch(10)
    assert e.value.code == capi.HS_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_bad_parameters_rejected():
    lib = capi.load()
    ctx = C.c_void_p()
    for bad in (capi.Params(0, 4, 4, 50.0, 30.0, 0, 0, 0, 0), capi.Params(10, 0, 4, 50.0, 30.0, 0, 0, 0, 0),
                capi.Params(10, 4, 999, 50.0, 30.0, 0, 0, 0, 0), capi.Params(33, 4, 4, 50.0, 30.0, 0, 0, 0, 0)):
        assert lib.hs_create(C.byref(ctx), 0, C.byref(bad)) == capi.HS_ERR_UNSUPPORTED
    bad = capi.Params(10, 4, 4, -1.0, 30.0, 0, 0, 0, 0)
    assert lib.hs_create(C.byref(ctx), 0, C.byref(bad)) == capi.HS_ERR_INVALID
    assert lib.hs_create(None, 0, None) == capi.HS_ERR_INVALID
    assert b"null" in lib.hs_last_error()


def test_evaluate_recall_null_arguments():
    # argument checks come before any device work, so they answer without a GPU
    lib = capi.load()
    r = capi.Recall()
    assert lib.hs_evaluate_recall(None, None, 0, None, 0, 1, C.byref(r)) == capi.HS_ERR_INVALID
    assert lib.hs_evaluate_recall_dev(None, None, 0, None, 0, 1, C.byref(r)) == capi.HS_ERR_INVALID
    assert b"null" in lib.hs_last_error()
    assert C.sizeof(capi.Recall) == 16 + 3 * 8 + 2 * 8 * capi.RECALL_BINS  # hs_recall of include/hsearch_b200.h


def test_hit_checksum_host_properties():
    rng = np.random.default_rng(5)
    hits = np.zeros(1000, dtype=capi.HIT_DTYPE)
    hits["query"] = rng.integers(0, 100, 1000)
    hits["table_first"] = rng.integers(0, 4, 1000)
    hits["db_id"] = rng.integers(0, 1 << 40, 1000)
    hits["dist2"] = rng.random(1000) * 900
    s = hb.HSearch.hits_checksum(hits)
    assert s == hb.HSearch.hits_checksum(hits[rng.permutation(1000)])
    parts = (hb.HSearch.hits_checksum(hits[:300]) + hb.HSearch.hits_checksum(hits[300:])) % (1 << 64)
    assert parts == s
    for field in ("query", "table_first", "db_id"):
        other = hits.copy()
        other[field][17] += 1
        assert hb.HSearch.hits_checksum(other) != s
    assert hb.HSearch.hits_checksum(hits[:0]) == 0


def test_expand_compact_hits_host():
    lib = capi.load()
    Q, id_bits = 5, 20
    counts = np.array([3, 0, 2, 0, 1])
    off = np.zeros(Q + 1, dtype=np.uint64)
    off[1:] = np.cumsum(counts)
    ids = np.array([5, 9, 7, 1, 2, 1048575], dtype=np.uint32)
    tabs = np.array([0, 0, 2, 1, 3, 3], dtype=np.uint32)
    idt = (ids | (tabs << id_bits)).astype(np.uint32)
    d2 = np.array([0.0, 1.5, 2.5, 3.5, 4.5, 899.999])
    ch = capi.CompactHits(capi.ptr(off, C.c_uint64), capi.ptr(idt, C.c_uint32), capi.ptr(d2, C.c_double), 6, id_bits)
    out = np.zeros(6, dtype=capi.HIT_DTYPE)
    capi.check(lib.hs_expand_hits(C.byref(ch), Q, 10 ** 10, out.ctypes.data))
    assert out["query"].tolist() == [0, 0, 0, 2, 2, 4]
    assert out["table_first"].tolist() == tabs.tolist()
    assert out["db_id"].tolist() == (ids.astype(np.uint64) + 10 ** 10).tolist()
    assert out["dist2"].tolist() == d2.tolist()
    assert lib.hs_expand_hits(None, Q, 0, None) == capi.HS_ERR_INVALID


def test_fragment_name_host():
    # name#i$j@KMER*cnt, the name being the first token of the header (protein2datapoints.cpp:61-65)
    assert hb.index.fragment_name("sp|P1|X some text", 3, 77, "ARNDC", 12) == "sp|P1|X#3$77@ARNDC*12"
    assert hb.index.fragment_name("  \tlead", 0, 0, "A", 0) == "lead#0$0@A*0"
    assert hb.index.fragment_name("", 4294967295, 5, "WYV", 2 ** 40) == "#4294967295$5@WYV*%d" % 2 ** 40
    lib = capi.load()
    buf = C.create_string_buffer(8)
    assert lib.hs_fragment_name(b"name", 1, 2, b"ARN", 3, 4, buf, 8) == capi.HS_ERR_CAPACITY
    assert lib.hs_fragment_name(None, 1, 2, b"ARN", 3, 4, buf, 8) == capi.HS_ERR_INVALID


def test_blosum_filter_embedding_contracts():
    """The table the tensor filter uses for the integer metric must contract: |e(a) - e(b)|^2 <= D[a][b]
    for all residue pairs (then sum_p |e(x_p) - e(y_p)|^2 <= the integer window distance,
    evaluate_correlation.cpp:26-41, and the filter cannot lose a pair within R)."""
    lib = capi.load()
    e = np.zeros(160, dtype=np.float64)
    capi.check(lib.hs_get_blosum_filter_embedding(e.ctypes.data_as(C.POINTER(C.c_double))))
    D = np.zeros(400, dtype=np.int32)
    capi.check(lib.hs_get_blosum_metric(D.ctypes.data_as(C.POINTER(C.c_int32))))
    e = e.reshape(20, 8)
    D = D.reshape(20, 20)
    d2 = ((e[:, None, :] - e[None, :, :]) ** 2).sum(-1)
    assert np.all(d2 <= D * (1 - 1e-11) + 1e-15)
    off = ~np.eye(20, dtype=bool)
    ratio = d2[off] / D[off]
    assert ratio.max() > 0.999 and ratio.mean() > 0.5      # scaled until the tightest pair touches
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 20, size=(2, 200000, 10))
    assert np.all(d2[a, b].sum(1) <= D[a, b].sum(1))
