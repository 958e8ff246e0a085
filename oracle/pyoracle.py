"""ctypes front end of the parity oracle.  TEST INFRASTRUCTURE ONLY.

Loads oracle/liboracle.so (our C restatement, hs_oracle.c) and, when present,
oracle/_ref/libref_*.so (the reference's own sources compiled in place by
oracle/Makefile).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module; nothing under
hsearch_b200/ does.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class Hit(C.Structure):
    _fields_ = [("query", C.c_uint32), ("table_first", C.c_uint32), ("db_id", C.c_uint64),
                ("dist2", C.c_double)]


HIT_DTYPE = np.dtype([("query", "<u4"), ("table_first", "<u4"), ("db_id", "<u8"), ("dist2", "<f8")])


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "hs_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference/hclust"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


_dbl = C.POINTER(C.c_double)
_u8 = C.POINTER(C.c_uint8)
_u32 = C.POINTER(C.c_uint32)
_u64 = C.POINTER(C.c_uint64)
_i32 = C.POINTER(C.c_int)


class Oracle:
    def __init__(self):
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        self.L = L
        L.orc_lsh_generate.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, _dbl, _dbl]
        L.orc_klsh_generate.argtypes = [C.c_uint32, C.c_uint32, C.c_double, _dbl, _dbl, _dbl]
        L.orc_klsh_hash.argtypes = [_dbl, C.c_uint32, C.c_uint32, _dbl, _dbl, _dbl]
        L.orc_klsh_hash.restype = C.c_uint64
        L.orc_kmer3_features.argtypes = [C.c_char_p, C.c_uint32, _dbl]
        L.orc_hash_codes.argtypes = [_u8, C.c_uint64, C.c_uint32, _dbl, _dbl, _dbl, C.c_uint32,
                                     C.c_uint32, C.c_double, _i32]
        L.orc_hash_points.argtypes = [_dbl, C.c_uint64, C.c_uint32, _dbl, _dbl, C.c_uint32, C.c_uint32,
                                      C.c_double, _i32]
        L.orc_dist2.argtypes = [_dbl, _dbl, C.c_uint32]
        L.orc_dist2.restype = C.c_double
        L.orc_search.argtypes = [_dbl, C.c_uint64, _dbl, C.c_uint32, C.c_uint32, _dbl, _dbl, C.c_uint32,
                                 C.c_uint32, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_uint64,
                                 _u64, _u64]
        L.orc_search.restype = C.c_uint64
        L.orc_bruteforce.argtypes = [_dbl, C.c_uint64, _dbl, C.c_uint32, C.c_uint32, C.c_double, C.c_int,
                                     C.c_void_p, C.c_uint64]
        L.orc_bruteforce.restype = C.c_uint64
        L.orc_bruteforce_int.argtypes = [_u8, C.c_uint64, _u8, C.c_uint32, C.c_uint32, C.c_int,
                                         C.c_void_p, C.c_uint64]
        L.orc_bruteforce_int.restype = C.c_uint64
        L.orc_union_find_labels.argtypes = [C.c_uint32, _u32, _u32, C.c_uint64, _u32]
        L.orc_cluster.argtypes = [_u8, C.c_uint64, C.c_uint32, _dbl, _dbl, _dbl, C.c_uint32, C.c_uint32,
                                  C.c_double, C.c_double, C.c_int, _u32]
        L.orc_cluster.restype = C.c_uint64
        L.orc_extract_windows.argtypes = [_u8, _u32, C.c_uint32, C.c_uint32, C.c_uint32, _u8, _u32]
        L.orc_extract_windows.restype = C.c_uint64
        L.orc_protein_id.argtypes = [_u32, C.c_uint32, C.c_uint32]
        L.orc_protein_id.restype = C.c_uint32
        L.orc_proteindb_code.argtypes = [C.c_char]
        L.orc_proteindb_stored_letter.argtypes = [C.c_char]
        L.orc_proteindb_stored_letter.restype = C.c_char
        L.orc_orf6.argtypes = [C.c_char_p, C.c_int, C.c_char_p, _i32]
        L.orc_weight.argtypes = [C.c_double, C.c_double]
        L.orc_weight.restype = C.c_double
        L.orc_evaluate.argtypes = [C.c_void_p, _dbl, C.c_uint64, C.c_void_p, C.c_uint64, C.c_double, C.c_uint32, _dbl,
                                   _dbl, _u64, _u64, _u64, _u64, _u64]
        L.orc_distance_int.argtypes = [_u8, _u8, C.c_uint32, _i32]
        L.orc_similarity_int.argtypes = [_u8, _u8, C.c_uint32]
        L.orc_get_aa20.restype = C.c_char_p

    # ---- tables -----------------------------------------------------------
    def coordinates(self, print6=False):
        t = np.zeros(160, dtype=np.float64)
        (self.L.orc_get_coordinates_print6 if print6 else self.L.orc_get_coordinates)(_p(t, C.c_double))
        return t.reshape(20, 8)

    def blosum62(self):
        t = np.zeros(400, dtype=np.int32)
        self.L.orc_get_blosum62(_p(t, C.c_int))
        return t.reshape(20, 20)

    def blosum_metric(self):
        t = np.zeros(400, dtype=np.int32)
        self.L.orc_blosum_metric(_p(t, C.c_int))
        return t.reshape(20, 20)

    def triangle_violations(self, d):
        d = np.ascontiguousarray(d, dtype=np.int32)
        return self.L.orc_triangle_violations(_p(d, C.c_int))

    def base(self):
        t = np.zeros(26, dtype=np.int32)
        self.L.orc_get_base(_p(t, C.c_int))
        return t

    def aa20(self):
        return self.L.orc_get_aa20().decode()

    # ---- projection -------------------------------------------------------
    def lsh_generate(self, seed, dim, K, W):
        a = np.zeros((K, dim), dtype=np.float64)
        b = np.zeros(K, dtype=np.float64)
        self.L.orc_lsh_generate(seed, dim, K, W, _p(a, C.c_double), _p(b, C.c_double))
        return a, b

    def lsh_tables(self, seed_base, dim, K, L, W):
        """a [L][K][dim], b [L][K]; table l seeded seed_base + l."""
        ab = [self.lsh_generate(seed_base + l, dim, K, W) for l in range(L)]
        return np.stack([x[0] for x in ab]), np.stack([x[1] for x in ab])

    def klsh_generate(self, feat=512, bits=16, sigma=0.2):
        w = np.zeros((bits, feat)); t = np.zeros(bits); b = np.zeros(bits)
        self.L.orc_klsh_generate(feat, bits, sigma, _p(w, C.c_double), _p(t, C.c_double), _p(b, C.c_double))
        return w, t, b

    def klsh_hash(self, p, w, t, b):
        p = np.ascontiguousarray(p, dtype=np.float64)
        return int(self.L.orc_klsh_hash(_p(p, C.c_double), w.shape[1], w.shape[0], _p(w, C.c_double),
                                        _p(t, C.c_double), _p(b, C.c_double)))

    def kmer3_features(self, seq):
        f = np.zeros(512)
        s = seq.encode() if isinstance(seq, str) else seq
        self.L.orc_kmer3_features(s, len(s), _p(f, C.c_double))
        return f

    # ---- embedding / hash -------------------------------------------------
    def embed(self, codes, table):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        return np.ascontiguousarray(table[codes].reshape(codes.shape[0], -1))

    def hash_codes(self, codes, table, a, b, W):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        table = np.ascontiguousarray(table, dtype=np.float64)
        N, ln = codes.shape
        L, K, dim = a.shape
        out = np.zeros((N, L, K), dtype=np.int32)
        self.L.orc_hash_codes(_p(codes, C.c_uint8), N, ln, _p(table, C.c_double), _p(a, C.c_double),
                              _p(b, C.c_double), K, L, W, _p(out, C.c_int))
        return out

    def hash_points(self, pts, a, b, W):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        N, dim = pts.shape
        L, K, _ = a.shape
        out = np.zeros((N, L, K), dtype=np.int32)
        self.L.orc_hash_points(_p(pts, C.c_double), N, dim, _p(a, C.c_double), _p(b, C.c_double), K, L, W,
                               _p(out, C.c_int))
        return out

    @staticmethod
    def key_strings(buckets):
        """H4 (lsh.hpp:51-59): decimal concat, no separator.  buckets [..., K]."""
        flat = buckets.reshape(-1, buckets.shape[-1])
        return np.array(["".join(str(int(v)) for v in row) for row in flat], dtype=object).reshape(buckets.shape[:-1])

    # ---- search -----------------------------------------------------------
    def search(self, db, queries, a, b, W, R, pred=0, cap=None):
        db = np.ascontiguousarray(db, dtype=np.float64)
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        N, dim = db.shape
        Q = queries.shape[0]
        L, K, _ = a.shape
        cap = cap or max(1024, 64 * Q)
        while True:
            hits = np.zeros(cap, dtype=HIT_DTYPE)
            ts = np.zeros(L, dtype=np.uint64)
            nc = np.zeros(1, dtype=np.uint64)
            n = self.L.orc_search(_p(db, C.c_double), N, _p(queries, C.c_double), Q, dim, _p(a, C.c_double),
                                  _p(b, C.c_double), K, L, W, R, pred, hits.ctypes.data, cap, _p(ts, C.c_uint64),
                                  _p(nc, C.c_uint64))
            if n <= cap:
                return hits[:n], ts, int(nc[0])
            cap = int(n)

    def bruteforce(self, db, queries, R, pred=1, cap=None):
        db = np.ascontiguousarray(db, dtype=np.float64)
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        N, dim = db.shape
        Q = queries.shape[0]
        cap = cap or max(1024, 64 * Q)
        while True:
            hits = np.zeros(cap, dtype=HIT_DTYPE)
            n = self.L.orc_bruteforce(_p(db, C.c_double), N, _p(queries, C.c_double), Q, dim, R, pred,
                                      hits.ctypes.data, cap)
            if n <= cap:
                return hits[:n]
            cap = int(n)

    def bruteforce_int(self, db_codes, qcodes, R, cap=None):
        db_codes = np.ascontiguousarray(db_codes, dtype=np.uint8)
        N, ln = db_codes.shape
        if qcodes is not None:
            qcodes = np.ascontiguousarray(qcodes, dtype=np.uint8)
            Q = qcodes.shape[0]
            qp = _p(qcodes, C.c_uint8)
        else:
            Q = 0
            qp = None
        cap = cap or 1 << 16
        while True:
            hits = np.zeros(cap, dtype=HIT_DTYPE)
            n = self.L.orc_bruteforce_int(_p(db_codes, C.c_uint8), N, qp, Q, ln, int(R), hits.ctypes.data, cap)
            if n <= cap:
                return hits[:n]
            cap = int(n)

    def distance_int(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.uint8); y = np.ascontiguousarray(y, dtype=np.uint8)
        d = np.ascontiguousarray(self.blosum_metric().reshape(-1), dtype=np.int32)
        return self.L.orc_distance_int(_p(x, C.c_uint8), _p(y, C.c_uint8), len(x), _p(d, C.c_int))

    def similarity_int(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.uint8); y = np.ascontiguousarray(y, dtype=np.uint8)
        return self.L.orc_similarity_int(_p(x, C.c_uint8), _p(y, C.c_uint8), len(x))

    def dist2(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
        return self.L.orc_dist2(_p(x, C.c_double), _p(y, C.c_double), len(x))

    # ---- cluster ----------------------------------------------------------
    def union_find_labels(self, n, eu, ev):
        eu = np.ascontiguousarray(eu, dtype=np.uint32); ev = np.ascontiguousarray(ev, dtype=np.uint32)
        out = np.zeros(n, dtype=np.uint32)
        self.L.orc_union_find_labels(n, _p(eu, C.c_uint32), _p(ev, C.c_uint32), len(eu), _p(out, C.c_uint32))
        return out

    def cluster(self, codes, table, a, b, W, R, metric=0):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        table = np.ascontiguousarray(table, dtype=np.float64)
        N, ln = codes.shape
        L, K, _ = a.shape
        out = np.zeros(N, dtype=np.uint32)
        ne = self.L.orc_cluster(_p(codes, C.c_uint8), N, ln, _p(table, C.c_double), _p(a, C.c_double),
                                _p(b, C.c_double), K, L, W, R, metric, _p(out, C.c_uint32))
        return out, int(ne)

    # ---- sequence front end ------------------------------------------------
    def extract_windows(self, residues, start_index, L, stride=1):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        start_index = np.ascontiguousarray(start_index, dtype=np.uint32)
        nprot = len(start_index) - 1
        n = self.L.orc_extract_windows(_p(residues, C.c_uint8), _p(start_index, C.c_uint32), nprot, L, stride,
                                       None, None)
        codes = np.zeros((n, L), dtype=np.uint8)
        pos = np.zeros(n, dtype=np.uint32)
        self.L.orc_extract_windows(_p(residues, C.c_uint8), _p(start_index, C.c_uint32), nprot, L, stride,
                                   _p(codes, C.c_uint8), _p(pos, C.c_uint32))
        return codes, pos

    def protein_id(self, start_index, pos):
        start_index = np.ascontiguousarray(start_index, dtype=np.uint32)
        return self.L.orc_protein_id(_p(start_index, C.c_uint32), len(start_index), pos)

    def proteindb_code(self, letter):
        return self.L.orc_proteindb_code(letter.encode())

    def orf6(self, dna):
        s = dna.encode()
        n = len(s)
        stride = n // 3 + 2
        buf = C.create_string_buffer(6 * stride)
        kept = (C.c_int * 6)()
        self.L.orc_orf6(s, n, buf, kept)
        out = []
        for f in range(6):
            if kept[f]:
                out.append(buf.raw[f * stride:(f + 1) * stride].split(b"\0")[0].decode())
        return out

    def weight(self, dis, R):
        return self.L.orc_weight(dis, R)

    def evaluate(self, truth, found, R, dis=None, nbins=500):
        """evaulate() (motif_both_points.cpp:100-165) on HIT_DTYPE arrays; both are sorted by
        (query, db id) here, dis defaults to sqrt(dist2) of the truth list."""
        truth = np.sort(np.ascontiguousarray(truth, dtype=HIT_DTYPE), order=["query", "db_id"])
        found = np.sort(np.ascontiguousarray(found, dtype=HIT_DTYPE), order=["query", "db_id"])
        tdis = np.ascontiguousarray(np.sqrt(truth["dist2"]) if dis is None else dis, dtype=np.float64)
        tp, fn = C.c_double(), C.c_double()
        cnt = np.zeros(3, dtype=np.uint64)
        tpb, fnb = np.zeros(nbins, dtype=np.uint64), np.zeros(nbins, dtype=np.uint64)
        rc = self.L.orc_evaluate(truth.ctypes.data_as(C.c_void_p), _p(tdis, C.c_double), len(truth),
                                 found.ctypes.data_as(C.c_void_p), len(found), R, nbins, C.byref(tp), C.byref(fn),
                                 _p(cnt[0:1], C.c_uint64), _p(cnt[1:2], C.c_uint64), _p(cnt[2:3], C.c_uint64),
                                 _p(tpb, C.c_uint64), _p(fnb, C.c_uint64))
        if rc != 0:
            raise ValueError("err: a ground-truth distance exceeds R + 0.1 (the reference exits)")
        return {"tp": tp.value, "fn": fn.value, "recall": tp.value / (tp.value + fn.value) if tp.value + fn.value else float("nan"),
                "n_tp": int(cnt[0]), "n_fn": int(cnt[1]), "n_extra": int(cnt[2]), "tp_bin": tpb, "fn_bin": fnb}


class Reference:
    """The reference's own code (oracle/_ref), when it was built."""

    def __init__(self):
        self.search_lib = C.CDLL(os.path.join(REF_DIR, "libref_search.so"))
        self.nolsh_lib = C.CDLL(os.path.join(REF_DIR, "libref_nolsh.so"))
        self.pc_lib = C.CDLL(os.path.join(REF_DIR, "libref_pcluster.so"))
        S = self.search_lib
        S.ref_lsh_generate.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, _dbl, _dbl]
        S.ref_hash_points.argtypes = [_dbl, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                      C.c_uint64, _i32, C.c_char_p, C.c_uint32]
        S.ref_search.argtypes = [_dbl, C.c_uint64, _dbl, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.c_double, C.c_double, C.c_uint64, C.c_char_p, C.c_void_p, _dbl, C.c_uint64,
                                 _u64, _dbl]
        S.ref_search.restype = C.c_uint64
        S.ref_build_tables_seconds.argtypes = [_dbl, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                               C.c_double, C.c_uint64, _u64]
        S.ref_build_tables_seconds.restype = C.c_double
        S.ref_weight.argtypes = [C.c_double, C.c_double]
        S.ref_weight.restype = C.c_double
        S.ref_evaluate.argtypes = [C.c_char_p, C.c_char_p, C.c_double]
        S.ref_evaluate.restype = C.c_double
        B = self.nolsh_lib
        B.ref_bruteforce.argtypes = [_dbl, C.c_uint64, _dbl, C.c_uint32, C.c_uint32, C.c_double, C.c_char_p,
                                     C.c_void_p, _dbl, C.c_uint64, _dbl]
        B.ref_bruteforce.restype = C.c_uint64
        P = self.pc_lib
        P.ref_union_find.argtypes = [_u32, C.c_uint32, _u32, _u32, C.c_uint64, _u32]
        P.ref_klsh_hash.argtypes = [_dbl, C.c_uint32, C.c_uint32, C.c_double]
        P.ref_klsh_hash.restype = C.c_uint64
        P.ref_kmer2integer.argtypes = [C.c_char_p]
        P.ref_kmer2integer.restype = C.c_uint32
        P.ref_read_fasta.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
        P.ref_read_fasta.restype = C.c_uint32

    @staticmethod
    def available():
        return os.path.exists(os.path.join(REF_DIR, "libref_search.so"))

    def lsh_generate(self, seed, dim, K, W):
        a = np.zeros((K, dim)); b = np.zeros(K)
        self.search_lib.ref_lsh_generate(seed, dim, K, W, _p(a, C.c_double), _p(b, C.c_double))
        return a, b

    def hash_points(self, pts, K, L, W, seed_base, want_keys=True):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        N, dim = pts.shape
        out = np.zeros((N, L, K), dtype=np.int32)
        stride = 12 * K + 1
        buf = C.create_string_buffer(N * L * stride) if want_keys else None
        self.search_lib.ref_hash_points(_p(pts, C.c_double), N, dim, K, L, W, seed_base, _p(out, C.c_int), buf,
                                        stride)
        keys = None
        if want_keys:
            raw = buf.raw
            keys = np.array([raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode()
                             for i in range(N * L)], dtype=object).reshape(N, L)
        return out, keys

    def search(self, db, queries, K, L, W, R, seed_base, cap=None):
        db = np.ascontiguousarray(db, dtype=np.float64)
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        N, dim = db.shape
        Q = queries.shape[0]
        cap = cap or max(1 << 16, 256 * Q)
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "hits.txt").encode()
            while True:
                hits = np.zeros(cap, dtype=HIT_DTYPE)
                printed = np.zeros(cap)
                ts = np.zeros(L, dtype=np.uint64)
                sec = np.zeros(1)
                n = self.search_lib.ref_search(_p(db, C.c_double), N, _p(queries, C.c_double), Q, dim, K, L, W, R,
                                               seed_base, path, hits.ctypes.data, _p(printed, C.c_double), cap,
                                               _p(ts, C.c_uint64), _p(sec, C.c_double))
                if n <= cap:
                    return hits[:n], printed[:n], ts, float(sec[0])
                cap = int(n)

    def build_tables_seconds(self, db, K, L, W, seed_base):
        db = np.ascontiguousarray(db, dtype=np.float64)
        N, dim = db.shape
        ts = np.zeros(L, dtype=np.uint64)
        s = self.search_lib.ref_build_tables_seconds(_p(db, C.c_double), N, dim, K, L, W, seed_base,
                                                     _p(ts, C.c_uint64))
        return s, ts

    def bruteforce(self, db, queries, R, cap=None):
        db = np.ascontiguousarray(db, dtype=np.float64)
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        N, dim = db.shape
        Q = queries.shape[0]
        cap = cap or max(1 << 16, 256 * Q)
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "bf.txt").encode()
            while True:
                hits = np.zeros(cap, dtype=HIT_DTYPE)
                printed = np.zeros(cap)
                sec = np.zeros(1)
                n = self.nolsh_lib.ref_bruteforce(_p(db, C.c_double), N, _p(queries, C.c_double), Q, dim, R, path,
                                                  hits.ctypes.data, _p(printed, C.c_double), cap,
                                                  _p(sec, C.c_double))
                if n <= cap:
                    return hits[:n], printed[:n], float(sec[0])
                cap = int(n)

    def evaluate(self, truth, found, R):
        """The reference's evaulate() (motif_both_points.cpp:100-165) run on text files written from
        the two HIT_DTYPE lists: zero-padded names make its string order equal (query, db id),
        distances are written with 17 significant digits (read back exactly).  Returns
        (tp/(tp+fn), rows of <out>.accuracy.txt as token lists)."""
        truth = np.sort(np.ascontiguousarray(truth, dtype=HIT_DTYPE), order=["query", "db_id"])
        with tempfile.TemporaryDirectory() as d:
            gt, out = os.path.join(d, "gt.txt"), os.path.join(d, "out.txt")
            with open(gt, "w") as f:
                for h in truth:
                    f.write("m%09d p%012d %.17g\n" % (h["query"], h["db_id"], np.sqrt(h["dist2"])))
            with open(out, "w") as f:
                for h in found:
                    f.write("m%09d p%012d %.17g\n" % (h["query"], h["db_id"], np.sqrt(h["dist2"])))
            r = self.search_lib.ref_evaluate(gt.encode(), out.encode(), R)
            rows = [ln.split() for ln in open(out + ".accuracy.txt").read().splitlines() if ln.strip()]
        return float(r), rows

    def union_find_roots(self, ids, eu, ev):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        eu = np.ascontiguousarray(eu, dtype=np.uint32); ev = np.ascontiguousarray(ev, dtype=np.uint32)
        out = np.zeros(len(ids), dtype=np.uint32)
        self.pc_lib.ref_union_find(_p(ids, C.c_uint32), len(ids), _p(eu, C.c_uint32), _p(ev, C.c_uint32), len(eu),
                                   _p(out, C.c_uint32))
        return out

    def klsh_hash(self, p, bits=16, sigma=0.2):
        p = np.ascontiguousarray(p, dtype=np.float64)
        return int(self.pc_lib.ref_klsh_hash(_p(p, C.c_double), len(p), bits, sigma))

    def kmer2integer(self, kmer):
        return int(self.pc_lib.ref_kmer2integer(kmer.encode()))

    def read_fasta(self, path):
        buf = C.create_string_buffer(1 << 24)
        n = self.pc_lib.ref_read_fasta(path.encode(), buf, len(buf))
        recs = [ln.split("\t") for ln in buf.value.decode().split("\n") if ln]
        return n, recs
