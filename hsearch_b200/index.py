"""Host-side mirror of the reference's search/cluster drivers over the C ABI.

`HSearch` plays the role of the reference's `Search()` (hclust/src/hclust/
motif_both_points.cpp:195-250), its brute-force twin
(motif_both_points_noLSH.cpp:36-56) and the union-find cluster composition
(pcluster/src/pcluster/union_find.cpp:3-33): same parameter names (hash_K,
hash_L, hash_W, hash_R), same meaning, errors raised instead of exit codes.
All compute goes through libhsearch_b200.so; there is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import (HIT_DTYPE, HS_FLAG_HASH_AUDIT, HS_FLAG_HASH_EXACT, HS_FLAG_SORT_HITS, HS_METRIC_BLOSUM_INT,
                   HS_METRIC_EUCLID_FP64, HS_PRED_D2_LE_R2, HS_PRED_SQRT_LE_R, HS_TABLE_FULL, HS_TABLE_PRINT6,
                   HsError, Params, Stats, check, ptr)

AA_ORDER = "ARNDCQEGHILKMFPSTWYV"  # residue code i <-> letter (row order of `coordinates`, util.hpp:21-42)


def coordinates(table_variant=HS_TABLE_FULL):
    t = np.zeros(160, dtype=np.float64)
    check(capi.load().hs_get_coordinates(table_variant, ptr(t, C.c_double)))
    return t.reshape(20, 8)


def blosum_metric():
    t = np.zeros(400, dtype=np.int32)
    check(capi.load().hs_get_blosum_metric(ptr(t, C.c_int32)))
    return t.reshape(20, 20)


def generate_projection(seed_base, dim, hash_K, hash_L, hash_W):
    """LSH::LSH (lsh.hpp:10-31) for hash_L tables; table l is seeded seed_base + l.
    Returns a [L][K][dim], b [L][K]."""
    lib = capi.load()
    a = np.zeros((hash_L, hash_K, dim), dtype=np.float64)
    b = np.zeros((hash_L, hash_K), dtype=np.float64)
    for l in range(hash_L):
        check(lib.hs_generate_projection(seed_base + l, dim, hash_K, hash_W, ptr(a[l], C.c_double),
                                         ptr(b[l], C.c_double)))
    return a, b


def encode(seqs):
    """Letters -> residue codes (base[c-'A'], util.hpp:92)."""
    lib = capi.load()
    out = np.zeros((len(seqs), len(seqs[0]) if seqs else 0), dtype=np.uint8)
    for i, s in enumerate(seqs):
        for j, ch in enumerate(s):
            c = lib.hs_letter_to_code(ch.encode())
            if c < 0:
                raise ValueError(f"'{ch}' is not one of the 20 amino-acid letters")
            out[i, j] = c
    return out


def fragment_name(header, protein_index, offset, kmer, cnt):
    """`name#i$j@KMER*cnt` of protein2datapoints.cpp:61-65."""
    buf = C.create_string_buffer(len(header) + len(kmer) + 64)
    check(capi.load().hs_fragment_name(header.encode(), protein_index, offset, kmer.encode(), len(kmer), cnt, buf, len(buf)))
    return buf.value.decode()


def pack_key_string(s, key_words):
    w = np.zeros(key_words, dtype=np.uint64)
    check(capi.load().hs_pack_key_string(s.encode(), key_words, ptr(w, C.c_uint64)))
    return w


class HSearch:
    def __init__(self, kmer_length, hash_K=4, hash_L=4, hash_W=50.0, hash_R=200.0, table_variant=HS_TABLE_PRINT6,
                 metric=HS_METRIC_EUCLID_FP64, predicate=HS_PRED_D2_LE_R2, flags=HS_FLAG_SORT_HITS, device=0):
        self.lib = capi.load()
        self.params = Params(kmer_length, hash_K, hash_L, float(hash_W), float(hash_R), table_variant, metric,
                             predicate, flags)
        self.ctx = C.c_void_p()
        check(self.lib.hs_create(C.byref(self.ctx), device, C.byref(self.params)))
        self.len = kmer_length
        self.dim = 8 * kmer_length
        self.K, self.L = hash_K, hash_L
        self.device = device
        self.id_base = 0

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx.value:
            self.lib.hs_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- setup ---------------------------------------------------------------
    def set_projection(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert a.shape == (self.L, self.K, self.dim) and b.shape == (self.L, self.K)
        check(self.lib.hs_set_projection(self.ctx, ptr(a, C.c_double), ptr(b, C.c_double)))

    def seed_projection(self, seed_base):
        a, b = generate_projection(seed_base, self.dim, self.K, self.L, self.params.W)
        self.set_projection(a, b)
        return a, b

    def set_coordinates(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64).reshape(160)
        check(self.lib.hs_set_coordinates(self.ctx, ptr(t, C.c_double)))

    def load_fragments(self, codes, id_base=0):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        assert codes.ndim == 2 and codes.shape[1] == self.len
        check(self.lib.hs_load_fragments(self.ctx, ptr(codes, C.c_uint8), codes.shape[0], id_base))
        self.id_base = id_base

    def load_fragments_dev(self, dev_ptr, n, id_base=0):
        check(self.lib.hs_load_fragments_dev(self.ctx, C.c_void_p(dev_ptr), n, id_base))
        self.id_base = id_base

    def extract_windows(self, residues, start_index, stride=1, id_base=0, want_pos=True):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        start_index = np.ascontiguousarray(start_index, dtype=np.uint32)
        nprot = len(start_index) - 1
        nfrag = C.c_uint64(0)
        cap = max(1, len(residues)) if want_pos else 0
        pos = np.zeros(cap, dtype=np.uint32) if want_pos else None
        check(self.lib.hs_extract_windows(self.ctx, ptr(residues, C.c_uint8), ptr(start_index, C.c_uint32), nprot,
                                          stride, id_base, ptr(pos, C.c_uint32) if want_pos else None, cap,
                                          C.byref(nfrag)))
        self.id_base = id_base
        return nfrag.value, (pos[:nfrag.value] if want_pos else None)

    def protein_id(self, start_index, pos):
        """ProteinDB::ProteinID for an array of global residue positions (device binary search)."""
        start_index = np.ascontiguousarray(start_index, dtype=np.uint32)
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        out = np.zeros(len(pos), dtype=np.uint32)
        check(self.lib.hs_protein_id(self.ctx, ptr(start_index, C.c_uint32), len(start_index), ptr(pos, C.c_uint32),
                                     len(pos), ptr(out, C.c_uint32)))
        return out

    @property
    def num_fragments(self):
        return int(self.lib.hs_num_fragments(self.ctx))

    # ---- hash / index ----------------------------------------------------------
    def hash(self, want_buckets=False):
        n = self.num_fragments
        out = np.zeros((n, self.L, self.K), dtype=np.int32) if want_buckets else None
        check(self.lib.hs_hash(self.ctx, ptr(out, C.c_int32) if want_buckets else None))
        return out

    def hash_audit(self):
        """Residual FP32 boundary flips of the current keys against the all-FP64 hash (must be 0)."""
        n = C.c_uint64(0)
        check(self.lib.hs_hash_audit(self.ctx, C.byref(n)))
        return n.value

    def keys(self, table):
        kw = self.stats().key_words
        out = np.zeros((self.num_fragments, kw), dtype=np.uint64)
        check(self.lib.hs_get_keys(self.ctx, table, ptr(out, C.c_uint64)))
        return out

    def build_index(self):
        check(self.lib.hs_build_index(self.ctx))

    def table_sizes(self):
        s = np.zeros(self.L, dtype=np.uint64)
        check(self.lib.hs_table_sizes(self.ctx, ptr(s, C.c_uint64)))
        return s

    def table(self, l):
        n = self.num_fragments
        nb = int(self.table_sizes()[l])
        ids = np.zeros(n, dtype=np.uint32)
        starts = np.zeros(nb + 1, dtype=np.uint32)
        check(self.lib.hs_get_table(self.ctx, l, ptr(ids, C.c_uint32), ptr(starts, C.c_uint32)))
        return ids, starts

    def stats(self):
        s = Stats()
        check(self.lib.hs_get_stats(self.ctx, C.byref(s)))
        return s

    def stream_ptr(self):
        p = C.c_void_p()
        check(self.lib.hs_get_stream(self.ctx, C.byref(p)))
        return p.value or 0

    def search_points_dev(self, q_dev_ptr, Q, hits_dev_ptr, cap):
        """Device-resident queries and hit buffer; returns the hit count (may exceed cap)."""
        n = C.c_uint64(0)
        rc = self.lib.hs_search_points_dev(self.ctx, C.c_void_p(q_dev_ptr), Q, C.c_void_p(hits_dev_ptr), cap,
                                           C.byref(n))
        if rc != capi.HS_ERR_CAPACITY:
            check(rc)
        return n.value

    def bruteforce_points_dev(self, q_dev_ptr, Q, hits_dev_ptr, cap):
        """Brute force with device-resident queries / hit buffer; returns the hit count (may exceed cap)."""
        n = C.c_uint64(0)
        rc = self.lib.hs_bruteforce_points_dev(self.ctx, C.c_void_p(q_dev_ptr), Q, C.c_void_p(hits_dev_ptr), cap,
                                               C.byref(n))
        if rc != capi.HS_ERR_CAPACITY:
            check(rc)
        return n.value

    # ---- search ------------------------------------------------------------------
    def _call_hits(self, fn, qarr, qctype, Q, cap, grow=True):
        cap = int(cap)
        while True:
            hits = np.zeros(max(cap, 1), dtype=HIT_DTYPE)
            n = C.c_uint64(0)
            rc = fn(self.ctx, ptr(qarr, qctype) if qarr is not None else None, Q, hits.ctypes.data, cap, C.byref(n))
            if rc == capi.HS_ERR_CAPACITY and grow:
                cap = int(n.value)
                continue
            check(rc)
            return hits[:n.value]

    def search_points(self, qpoints, cap=1 << 20):
        q = np.ascontiguousarray(qpoints, dtype=np.float64).reshape(-1, self.dim)
        return self._call_hits(self.lib.hs_search_points, q, C.c_double, q.shape[0], cap)

    def search_codes(self, qcodes, cap=1 << 20):
        q = np.ascontiguousarray(qcodes, dtype=np.uint8).reshape(-1, self.len)
        return self._call_hits(self.lib.hs_search_codes, q, C.c_uint8, q.shape[0], cap)

    def search_points_compact(self, qpoints, cap=1 << 20, expand=True):
        """hs_search_points_compact: per-query CSR (offsets, id | table << id_bits, dist2); with
        expand=True the hs_hit records hs_search_points would return (hs_expand_hits)."""
        q = np.ascontiguousarray(qpoints, dtype=np.float64).reshape(-1, self.dim)
        Q = q.shape[0]
        cap = int(cap)
        while True:
            off = np.zeros(Q + 1, dtype=np.uint64)
            idt = np.zeros(max(cap, 1), dtype=np.uint32)
            d2 = np.zeros(max(cap, 1), dtype=np.float64)
            ch = capi.CompactHits(ptr(off, C.c_uint64), ptr(idt, C.c_uint32), ptr(d2, C.c_double), cap, 0)
            n = C.c_uint64(0)
            rc = self.lib.hs_search_points_compact(self.ctx, ptr(q, C.c_double), Q, C.byref(ch), C.byref(n))
            if rc == capi.HS_ERR_CAPACITY:
                cap = int(n.value)
                continue
            check(rc)
            break
        if not expand:
            return off, idt[:n.value], d2[:n.value], int(ch.id_bits)
        hits = np.zeros(max(int(n.value), 1), dtype=HIT_DTYPE)
        check(self.lib.hs_expand_hits(C.byref(ch), Q, self.id_base, hits.ctypes.data))
        return hits[:n.value]

    def hits_checksum_dev(self, hits_dev_ptr, n):
        s = C.c_uint64(0)
        check(self.lib.hs_hits_checksum_dev(self.ctx, C.c_void_p(hits_dev_ptr), n, C.byref(s)))
        return s.value

    def bruteforce_points(self, qpoints, cap=1 << 20):
        q = np.ascontiguousarray(qpoints, dtype=np.float64).reshape(-1, self.dim)
        return self._call_hits(self.lib.hs_bruteforce_points, q, C.c_double, q.shape[0], cap)

    def bruteforce_codes(self, qcodes=None, cap=1 << 20):
        if qcodes is None:
            return self._call_hits(self.lib.hs_bruteforce_codes, None, C.c_uint8, 0, cap)
        q = np.ascontiguousarray(qcodes, dtype=np.uint8).reshape(-1, self.len)
        return self._call_hits(self.lib.hs_bruteforce_codes, q, C.c_uint8, q.shape[0], cap)

    @staticmethod
    def hits_checksum(hits):
        """Order-independent checksum of a HIT_DTYPE array (hs_hits_checksum, host)."""
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        s = C.c_uint64(0)
        check(capi.load().hs_hits_checksum(hits.ctypes.data_as(C.c_void_p), len(hits), C.byref(s)))
        return s.value

    @staticmethod
    def _recall_dict(r):
        tpb = np.frombuffer(r.tp_bin, dtype=np.uint64).copy()
        fnb = np.frombuffer(r.fn_bin, dtype=np.uint64).copy()
        return {"tp": r.tp, "fn": r.fn, "recall": r.tp / (r.tp + r.fn) if (r.tp + r.fn) else float("nan"),
                "n_tp": int(r.n_tp), "n_fn": int(r.n_fn), "n_extra": int(r.n_extra), "tp_bin": tpb, "fn_bin": fnb}

    def evaluate_recall(self, truth, found, Q):
        """evaulate() of motif_both_points.cpp:100-165 on binary hit lists (HIT_DTYPE arrays);
        `found` in the order the search returns with HS_FLAG_SORT_HITS."""
        truth = np.ascontiguousarray(truth, dtype=HIT_DTYPE)
        found = np.ascontiguousarray(found, dtype=HIT_DTYPE)
        r = capi.Recall()
        check(self.lib.hs_evaluate_recall(self.ctx, truth.ctypes.data_as(C.c_void_p), len(truth),
                                          found.ctypes.data_as(C.c_void_p), len(found), Q, C.byref(r)))
        return self._recall_dict(r)

    def evaluate_recall_dev(self, truth_ptr, n_truth, found_ptr, n_found, Q):
        """Same with both lists resident on the device (raw device pointers)."""
        r = capi.Recall()
        check(self.lib.hs_evaluate_recall_dev(self.ctx, C.c_void_p(truth_ptr), n_truth, C.c_void_p(found_ptr), n_found,
                                              Q, C.byref(r)))
        return self._recall_dict(r)

    def cluster(self):
        out = np.zeros(self.num_fragments, dtype=np.uint32)
        check(self.lib.hs_cluster(self.ctx, ptr(out, C.c_uint32)))
        return out

    def union_find(self, n, eu, ev):
        """Connected components of an explicit edge list over ids 0..n-1 (label = smallest id)."""
        eu = np.ascontiguousarray(eu, dtype=np.uint32)
        ev = np.ascontiguousarray(ev, dtype=np.uint32)
        out = np.zeros(n, dtype=np.uint32)
        check(self.lib.hs_union_find(self.ctx, n, ptr(eu, C.c_uint32), ptr(ev, C.c_uint32), len(eu), ptr(out, C.c_uint32)))
        return out

    def greedy_cluster(self):
        """hclust2's Clustering(): (centre of every fragment, round it joined in, merged[] flags)."""
        n = self.num_fragments
        center = np.zeros(n, dtype=np.uint32)
        rnd = np.zeros(n, dtype=np.uint32)
        state = np.zeros(n, dtype=np.uint8)
        check(self.lib.hs_greedy_cluster(self.ctx, ptr(center, C.c_uint32), ptr(rnd, C.c_uint32), ptr(state, C.c_uint8)))
        return center, rnd, state

    # ---- sequence front ends ------------------------------------------------
    @staticmethod
    def _concat(seqs):
        data = b"".join(s.encode() if isinstance(s, str) else s for s in seqs)
        start = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum([len(s) for s in seqs], out=start[1:])
        return data, start

    def kmer3_klsh(self, proteins, w, t, b, want_features=False):
        """3-mer histograms + KLSH values of a list of protein strings (pcluster PreClustering)."""
        data, start = self._concat(proteins)
        n = len(proteins)
        bits = w.shape[0]
        w = np.ascontiguousarray(w, dtype=np.float64)
        t = np.ascontiguousarray(t, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        feat = np.zeros((n, 512), dtype=np.uint32) if want_features else None
        hv = np.zeros(n, dtype=np.uint64)
        valid = np.zeros(n, dtype=np.uint8)
        fixed = C.c_uint64(0)
        check(self.lib.hs_kmer3_klsh(self.ctx, data, ptr(start, C.c_uint64), n, ptr(w, C.c_double), ptr(t, C.c_double),
                                     ptr(b, C.c_double), bits, ptr(feat, C.c_uint32) if want_features else None,
                                     ptr(hv, C.c_uint64), ptr(valid, C.c_uint8), C.byref(fixed)))
        return hv, valid, feat, fixed.value

    def orf6(self, dnas):
        """Six-frame translation of a list of DNA strings: per sequence the list of the six
        translated frames (the reference keeps those of length >= 6)."""
        data, start = self._concat(dnas)
        n = len(dnas)
        cap = 2 * int(start[-1]) + 6 * n
        out = C.create_string_buffer(max(cap, 1))
        ln = np.zeros((n, 6), dtype=np.int32)
        check(self.lib.hs_orf6(self.ctx, data, ptr(start, C.c_uint64), n, out, cap, ptr(ln, C.c_int32)))
        raw = out.raw
        res = []
        for s in range(n):
            ls = int(start[s + 1] - start[s])
            base = 2 * int(start[s]) + 6 * s
            res.append([raw[base + f * (ls // 3 + 1): base + f * (ls // 3 + 1) + int(ln[s, f])].decode() for f in range(6)])
        return res
