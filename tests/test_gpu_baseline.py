"""GPU parity at the sizes BASELINE.json names, plus the result formats and checks that
carry that parity to the full-size runs:

* C1 (configs[0]) exactly as SURVEY.md 8d states it -- 1 M fragments x 1 k queries, len 10,
  K = L = 4, W in {20, 50}, R = 30, table seeds 12345 + l -- against the reference's own
  Search() (hclust/src/hclust/motif_both_points.cpp:195-250, compiled in place under
  oracle/_ref) for hits, order and FP64 distances, and against the oracle restatement for the
  first-table column the reference does not print;
* the compact (CSR) hit layout expands to the very bytes hs_search_points returns;
* the order-independent hit checksum: device == host, and the checksums of the shards of a
  split database add up to the checksum of the whole (what an N-GPU run is compared by);
* the argument checks the advisor asked for (residue codes >= 20, context reuse)."""
import numpy as np
import pytest

import hsearch_b200 as hb
from tests.util import hits_as_tuples, planted_queries, random_codes

pytestmark = pytest.mark.gpu


def make(length, K, L, W, R, seed=12345, **kw):
    h = hb.HSearch(length, K, L, W, R, **kw)
    a, b = h.seed_projection(seed)
    return h, a, b


@pytest.mark.parametrize("W", [20.0, 50.0])
def test_c1_one_million_fragments_bit_exact_vs_reference(oracle, reference, W):
    n, q, length, K, L, R = 1_000_000, 1000, 10, 4, 4, 30.0
    codes = random_codes(n, length, seed=1)
    qcodes = planted_queries(codes, q, seed=2, frac=0.1)
    tab = oracle.coordinates(True)   # the 6-digit table motif_both_points really hashes
    db, qp = oracle.embed(codes, tab), oracle.embed(qcodes, tab)
    h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS | hb.HS_FLAG_HASH_AUDIT)
    h.load_fragments(codes)
    h.hash()
    st = h.stats()
    assert st.residual_flips == 0            # no FP32 boundary flip survives the guard band
    h.build_index()
    got = h.search_points(qp, cap=1 << 18)
    ref, printed, ts, _ = reference.search(db, qp, K, L, W, R, 12345, cap=1 << 18)
    assert np.array_equal(h.table_sizes(), ts)
    assert len(ref) > 100
    # the reference's output order, ids and FP64 distances (it does not print the first table)
    assert np.array_equal(got["query"], ref["query"])
    assert np.array_equal(got["db_id"], ref["db_id"])
    assert np.array_equal(got["dist2"], ref["dist2"])
    # the first-table column against the restatement (which equals the reference on the rest)
    want, ts2, _ = oracle.search(db, qp, a, b, W, R, pred=0, cap=1 << 18)
    assert np.array_equal(ts2, ts)
    assert np.array_equal(got, want.astype(got.dtype))
    # same search through the compact layout
    assert np.array_equal(h.search_points_compact(qp, cap=1 << 18), got)
    h.close()


@pytest.mark.parametrize("nq", [0, 1, 7, 300, 2600])
def test_compact_hits_expand_to_hs_hit(oracle, nq):
    length, K, L, W, R = 10, 4, 4, 50.0, 30.0
    codes = random_codes(40000, length, seed=21)
    tab = oracle.coordinates(True)
    qp = oracle.embed(planted_queries(codes, nq, seed=22, frac=0.5), tab) if nq else np.zeros((0, 80))
    h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
    h.load_fragments(codes, id_base=5_000_000_000)        # ids beyond 32 bits: only the local part is packed
    h.build_index()
    full = h.search_points(qp)
    comp = h.search_points_compact(qp)
    assert np.array_equal(full, comp)
    assert np.array_equal(h.search_points_compact(qp, cap=max(1, len(full) // 3)), full)   # capacity protocol
    off, idt, d2, id_bits = h.search_points_compact(qp, expand=False)
    assert off[0] == 0 and off[-1] == len(full) and np.all(np.diff(off.astype(np.int64)) >= 0)
    assert np.array_equal(np.repeat(np.arange(nq, dtype=np.uint32), np.diff(off.astype(np.int64))), full["query"])
    assert np.array_equal(idt >> id_bits, full["table_first"])
    if nq >= 300:
        assert len(full) > 100
        want, _, _ = oracle.search(oracle.embed(codes, tab), qp[:100], a, b, W, R)
        sel = full[full["query"] < 100].copy()
        sel["db_id"] -= 5_000_000_000
        assert hits_as_tuples(sel) == hits_as_tuples(want)
    h.close()


def test_hit_checksum_device_host_and_shards(oracle):
    """The checksum an N-GPU run is compared by: shards of the database, searched separately
    with their id_base, give hit lists whose checksums add up to the one-GPU list's."""
    import torch
    length, K, L, W, R = 10, 4, 4, 50.0, 30.0
    n = 60000
    codes = random_codes(n, length, seed=51)
    tab = oracle.coordinates(True)
    qp = oracle.embed(planted_queries(codes, 400, seed=52, frac=0.5), tab)
    h, a, b = make(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
    h.load_fragments(codes)
    h.build_index()
    whole = h.search_points(qp)
    assert len(whole) > 500
    s_host = hb.HSearch.hits_checksum(whole)
    dev = torch.from_numpy(whole.view(np.uint8).copy()).cuda()
    assert h.hits_checksum_dev(dev.data_ptr(), len(whole)) == s_host
    rng = np.random.default_rng(3)
    assert hb.HSearch.hits_checksum(whole[rng.permutation(len(whole))]) == s_host      # order does not matter
    broken = whole.copy()
    broken["dist2"][len(broken) // 2] = np.nextafter(broken["dist2"][len(broken) // 2], np.inf)
    assert hb.HSearch.hits_checksum(broken) != s_host                                   # one ulp of one distance does
    parts, total = [], 0
    for lo, hi in ((0, 25000), (25000, 25001), (25001, n)):
        h.load_fragments(codes[lo:hi], id_base=lo)
        h.build_index()
        part = h.search_points(qp)
        parts.append(part)
        total = (total + hb.HSearch.hits_checksum(part)) % (1 << 64)
    assert total == s_host
    merged = np.sort(np.concatenate(parts), order=["query", "table_first", "db_id"])
    assert np.array_equal(merged, whole)
    h.close()


def test_residue_codes_out_of_range_are_rejected():
    import ctypes as C
    from hsearch_b200 import capi
    h, a, b = make(10, 4, 4, 50.0, 30.0)
    codes = random_codes(5000, 10, seed=61)
    bad = codes.copy()
    bad[4321, 7] = 20
    for arr in (bad, np.where(codes == 3, 255, codes).astype(np.uint8)):
        with pytest.raises(hb.HsError) as e:
            h.load_fragments(arr)
        assert e.value.code == capi.HS_ERR_INVALID and "0..19" in str(e.value)
        assert h.num_fragments == 0
    # window extraction over residues holding hs_letter_to_code's -1 (0xff)
    res = random_codes(1, 400, seed=62).reshape(-1)
    res[123] = 0xFF
    with pytest.raises(hb.HsError):
        h.extract_windows(res, np.array([0, 400], dtype=np.uint32))
    h.load_fragments(codes)     # the context stays usable
    h.build_index()
    assert h.num_fragments == 5000
    h.close()


@pytest.mark.parametrize("no_rank", ["1", "0"])
def test_context_reuse_load_then_extract(oracle, monkeypatch, no_rank):
    """A reused context: records built for the first database must not serve the second (the
    packed-key path builds them lazily)."""
    monkeypatch.setenv("HS_NO_RANK", no_rank)
    length, K, L, W, R = 10, 4, 4, 50.0, 30.0
    tab = oracle.coordinates(True)
    h, a, b = make(length, K, L, W, R)
    codes = random_codes(20000, length, seed=71)
    h.load_fragments(codes)
    h.build_index()
    q = planted_queries(codes, 50, seed=72)
    want, _, _ = oracle.search(oracle.embed(codes, tab), oracle.embed(q, tab), a, b, W, R)
    assert hits_as_tuples(h.search_codes(q)) == hits_as_tuples(want)
    # second database through the window extractor: more fragments than the first
    res = random_codes(1, 30000, seed=73).reshape(-1)
    start = np.array([0, 9000, 9004, 30000], dtype=np.uint32)
    nfrag, pos = h.extract_windows(res, start)
    frags = np.stack([res[p:p + length] for p in pos])
    assert nfrag == len(frags) > 20000
    h.build_index()
    q2 = planted_queries(frags, 50, seed=74)
    want2, _, _ = oracle.search(oracle.embed(frags, tab), oracle.embed(q2, tab), a, b, W, R)
    got2 = h.search_codes(q2)
    assert len(want2) > 0 and hits_as_tuples(got2) == hits_as_tuples(want2)
    h.close()


def test_protein_id_of_hits(oracle):
    """ProteinDB::ProteinID (protein.hpp:28-39) on the device: window start positions -> proteins, so
    that a hit (fragment id) maps back to its protein; the end sentinel and positions past the end
    behave like the reference's search over all of start_index."""
    rng = np.random.default_rng(8)
    lens = rng.integers(1, 300, size=2000)
    lens[::97] = 0                                     # empty proteins share a start with their successor
    start = np.concatenate(([0], np.cumsum(lens))).astype(np.uint32)
    h = hb.HSearch(10, 4, 4, 50.0, 30.0)
    pos = np.concatenate((rng.integers(0, int(start[-1]) + 50, size=100000), start, start[1:] - 1)).astype(np.uint32)
    got = h.protein_id(start, pos)
    want = np.array([oracle.protein_id(start, int(p)) for p in pos[:3000]], dtype=np.uint32)
    assert np.array_equal(got[:3000], want)
    ref = np.searchsorted(start, pos, side="right").astype(np.int64) - 1     # last l with start[l] <= pos
    assert np.array_equal(got.astype(np.int64), ref)
    # through the window extractor: fragment -> protein -> data-point name
    res = random_codes(1, int(start[-1]), seed=9).reshape(-1)
    nfrag, fpos = h.extract_windows(res, start, stride=7)
    prot = h.protein_id(start, fpos)
    assert nfrag > 1000
    assert np.all(fpos >= start[prot]) and np.all(fpos + 10 <= start[prot + 1])
    kmer = "".join(hb.AA_ORDER[c] for c in res[fpos[5]:fpos[5] + 10])
    assert hb.index.fragment_name("p%d desc" % prot[5], int(prot[5]), int(fpos[5] - start[prot[5]]), kmer, 5) == \
        "p%d#%d$%d@%s*5" % (prot[5], prot[5], fpos[5] - start[prot[5]], kmer)
    h.close()
