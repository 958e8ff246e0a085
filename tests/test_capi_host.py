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
    assert C.sizeof(capi.Stats) == 9 * 8 + 4 * 4 + 15 * 4 + 4 + 8  # 15 floats, padded to 8, one more u64


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
