"""Sequence front ends: FASTA parsing, 3-mer features,
KLSH values, six-frame translation -- against the oracle and, where oracle/_ref is
present, against the reference's own code."""
import ctypes as C

import numpy as np
import pytest

import hsearch_b200 as hb
from hsearch_b200 import capi
from oracle.pyoracle import Reference

AA20 = "ARNDCEQGHILKMFPSTWYV"


def parse_fasta(text, ctx=None, seed=None):
    """hs_parse_fasta, or hs_parse_fasta_gpu on the context ctx; seed: srand() before each call."""
    lib = capi.load()
    libc = C.CDLL(None)
    data = text if isinstance(text, bytes) else text.encode()
    nseq, nnames, nres = C.c_uint32(), C.c_uint32(), C.c_uint64()
    if ctx is not None:
        def fn(*a):
            return lib.hs_parse_fasta_gpu(ctx, *a)
    else:
        fn = lib.hs_parse_fasta
    if seed is not None:
        libc.srand(seed)
    rc = fn(data, len(data), None, 0, None, 0, None, None, 0, C.byref(nseq), C.byref(nnames), C.byref(nres))
    assert rc in (capi.HS_OK, -3)
    if seed is not None:
        libc.srand(seed)
    res = C.create_string_buffer(max(1, nres.value))
    start = np.zeros(nseq.value + 1, dtype=np.uint64)
    nb = np.zeros(max(1, nnames.value), dtype=np.uint64)
    nl = np.zeros(max(1, nnames.value), dtype=np.uint32)
    capi.check(fn(data, len(data), res, nres.value, capi.ptr(start, C.c_uint64), len(start),
                  capi.ptr(nb, C.c_uint64), capi.ptr(nl, C.c_uint32), nnames.value, C.byref(nseq),
                  C.byref(nnames), C.byref(nres)))
    seqs = [res.raw[int(start[i]):int(start[i + 1])].decode("latin-1") for i in range(nseq.value)]
    names = [data[int(nb[i]):int(nb[i]) + int(nl[i])].decode("latin-1") for i in range(nnames.value)]
    return names, seqs


def test_parse_fasta_known_answers():
    text = ">p1 first protein\nARND\nCEQG\n\n>p2\nHILK\n>empty header only\n>p4 x\nMFPSTWYV\n"
    names, seqs = parse_fasta(text)
    # a name per header, a sequence only when non-empty (read_proteins.cpp:14-25)
    assert names == ["p1", "p2", "empty", "p4"]
    assert seqs == ["ARNDCEQG", "HILK", "MFPSTWYV"]
    names, seqs = parse_fasta("")
    assert names == [] and seqs == []
    names, seqs = parse_fasta(">only\n")
    assert names == ["only"] and seqs == []
    names, seqs = parse_fasta("ARND\n>x\n12 3*-\nAR ND\n")  # digits / punctuation dropped, no trailing header needed
    assert names == ["x"] and seqs == ["ARND", "ARND"]


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built")
def test_parse_fasta_matches_reference(tmp_path):
    rng = np.random.default_rng(3)
    lines = []
    for i in range(200):
        lines.append(f">prot{i} desc {i}" if i % 3 else f">prot{i}")
        n = int(rng.integers(0, 150))
        seq = "".join(AA20[c] for c in rng.integers(0, 20, size=n))
        for j in range(0, n, 60):
            lines.append(seq[j:j + 60])
    text = "\n".join(lines) + "\n"
    path = tmp_path / "p.fa"
    path.write_text(text)
    n, recs = Reference().read_fasta(str(path))
    names, seqs = parse_fasta(text)
    assert n == len(seqs)
    assert [r[1] for r in recs] == seqs
    # the reference pairs names and sequences by index (names of empty records shift the pairing)
    assert [r[0] for r in recs] == names[:len(seqs)]


def test_klsh_generate_matches_oracle(oracle):
    lib = capi.load()
    w = np.zeros((16, 512)); t = np.zeros(16); b = np.zeros(16)
    capi.check(lib.hs_klsh_generate(512, 16, 0.2, capi.ptr(w, C.c_double), capi.ptr(t, C.c_double), capi.ptr(b, C.c_double)))
    ow, ot, ob = oracle.klsh_generate(512, 16, 0.2)
    assert np.array_equal(w, ow) and np.array_equal(t, ot) and np.array_equal(b, ob)


def random_proteins(n, seed, lo=0, hi=400):
    rng = np.random.default_rng(seed)
    return ["".join(AA20[c] for c in rng.integers(0, 20, size=int(rng.integers(lo, hi)))) for _ in range(n)]


@pytest.mark.gpu
def test_kmer3_klsh_matches_oracle(oracle):
    prots = random_proteins(600, seed=4) + ["", "AR", "ARN", "A" * 1000, "ARNDCEQGHILKMFPSTWYV" * 40]
    w, t, b = oracle.klsh_generate(512, 16, 0.2)
    h = hb.HSearch(10, 4, 4, 50.0, 30.0)
    hv, valid, feat, fixed = h.kmer3_klsh(prots, w, t, b, want_features=True)
    for i, p in enumerate(prots):
        f = oracle.kmer3_features(p)
        assert np.array_equal(feat[i], f.astype(np.uint32)), i
        assert valid[i] == (1 if len(p) >= 3 else 0)
        if len(p) >= 3:
            assert int(hv[i]) == oracle.klsh_hash(f, w, t, b), i
    # SURVEY 8c known answers (reference KLSH(512,16,0.2)): the all-ones vector is not a 3-mer histogram,
    # but p[3]=2, p[77]=1, p[500]=5 is reachable only through the oracle; check the oracle pin here
    p = np.zeros(512); p[3] = 2; p[77] = 1; p[500] = 5
    assert oracle.klsh_hash(p, w, t, b) == 17156
    assert oracle.klsh_hash(np.ones(512), w, t, b) == 21252
    with pytest.raises(hb.HsError):
        h.kmer3_klsh(["ARNBX"], w, t, b)
    h.close()


@pytest.mark.gpu
@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built")
def test_kmer3_klsh_matches_reference(oracle):
    prots = random_proteins(100, seed=5, lo=3, hi=300)
    w, t, b = oracle.klsh_generate(512, 16, 0.2)
    r = Reference()
    h = hb.HSearch(10, 4, 4, 50.0, 30.0)
    hv, valid, _, _ = h.kmer3_klsh(prots, w, t, b)
    for i, p in enumerate(prots):
        assert int(hv[i]) == r.klsh_hash(oracle.kmer3_features(p)), i
    h.close()


@pytest.mark.gpu
def test_orf6_matches_oracle(oracle):
    rng = np.random.default_rng(6)
    dnas = ["".join("ACGT"[c] for c in rng.integers(0, 4, size=int(rng.integers(0, 500)))) for _ in range(300)]
    dnas += ["", "A", "ACG", "ATGAAACCCGGGTTTAAACCCGGGTTT", "TTT" * 50, "TAA" * 10]
    h = hb.HSearch(10, 4, 4, 50.0, 30.0)
    got = h.orf6(dnas)
    for i, d in enumerate(dnas):
        want = oracle.orf6(d)                       # frames the reference keeps (>= 6 residues), in frame order
        assert [f for f in got[i] if len(f) >= 6] == want, i
    with pytest.raises(hb.HsError):
        h.orf6(["ACGTNACGT"])
    h.close()


def test_parse_fasta_capacity_protocol():
    lib = capi.load()
    data = b">a\nARND\n>b\nCEQG\n"
    nseq, nnames, nres = C.c_uint32(), C.c_uint32(), C.c_uint64()
    res = C.create_string_buffer(3)   # too small: 8 residues
    start = np.zeros(3, dtype=np.uint64)
    rc = lib.hs_parse_fasta(data, len(data), res, 3, capi.ptr(start, C.c_uint64), 3, None, None, 0,
                            C.byref(nseq), C.byref(nnames), C.byref(nres))
    assert rc == capi.HS_ERR_CAPACITY and (nseq.value, nnames.value, nres.value) == (2, 2, 8)
    assert lib.hs_parse_fasta(None, 0, None, 0, None, 0, None, None, 0, None, None, None) == capi.HS_ERR_INVALID


@pytest.mark.gpu
def test_sequence_entry_points_edge_cases(oracle):
    w, t, b = oracle.klsh_generate(512, 16, 0.2)
    h = hb.HSearch(10, 4, 4, 50.0, 30.0)
    hv, valid, feat, fixed = h.kmer3_klsh([], w, t, b)
    assert len(hv) == 0 and fixed == 0
    assert h.orf6([]) == []
    # capacity error of hs_orf6
    data = b"ACGTACGTACGT"
    start = np.array([0, 12], dtype=np.uint64)
    ln = np.zeros(6, dtype=np.int32)
    out = C.create_string_buffer(8)
    assert h.lib.hs_orf6(h.ctx, data, capi.ptr(start, C.c_uint64), 1, out, 8, capi.ptr(ln, C.c_int32)) == capi.HS_ERR_CAPACITY
    h.close()


def _fasta_cases():
    rng = np.random.default_rng(11)
    cases = [b"", b">only\n", b"ARND\n>x\n12 3*-\nAR ND\n", b">a\n>b\n>c\nAR\n", b"\n\n>p q r\r\nARND\r\nbjz\n\n",
             b">no newline at the end", b"ARNDX", b">h\n" + b"ARNDCEQGHI" * 2000 + b"\n>t\nxyz\n",   # a line longer than 4 chunks
             b">" + b"n" * 9000 + b" tail\nAR\n",                                                      # a header longer than 2 chunks
             b"A\x00R\xffN\n>x\x00y\n"]
    alphabet = np.frombuffer(b"ARNDCEQGHILKMFPSTWYVBJOUXZarndxbz*- 0123456789>\r", dtype=np.uint8)
    for nlines, maxlen in ((4000, 80), (300, 5000), (30000, 12)):
        parts = []
        for i in range(nlines):
            r = rng.random()
            if r < 0.15:
                parts.append(b">prot%d" % i + (b" desc %d" % i if i % 3 else b""))
            elif r < 0.2:
                parts.append(b"")
            else:
                parts.append(alphabet[rng.integers(0, len(alphabet), size=int(rng.integers(0, maxlen)))].tobytes())
        cases.append(b"\n".join(parts) + (b"\n" if nlines % 2 else b""))
    return cases


@pytest.mark.gpu
def test_parse_fasta_gpu_equals_host_parser():
    """hs_parse_fasta_gpu (byte-parallel, fasta.cu) against hs_parse_fasta (ReadFASTAFile,
    read_proteins.cpp:6-41) under the same srand: names, sequences and replaced letters."""
    h = hb.HSearch(10, 4, 4, 50.0, 30.0)
    libc = C.CDLL(None)
    for k, text in enumerate(_fasta_cases()):
        want = parse_fasta(text, seed=100 + k)
        after_host = libc.rand()
        got = parse_fasta(text, ctx=h.ctx, seed=100 + k)
        after_dev = libc.rand()
        assert got == want, k
        assert after_host == after_dev     # the same number of rand() draws
    h.close()
