"""The drop-in command-line programs (hsearch_b200/bin/*, host C++ over the C ABI)
against the reference's own programs compiled in place (oracle/_ref/bin/*):
same argv grammar, same messages and exit codes, same output files.

CPU part: option parsing / help / missing-argument protocol
(smithlab OptionParser semantics, hclust/src/smithlab_cpp/OptionParser.cpp:156-184;
motif_both_points.cpp:324-335) and protein2datapoints (host-only program).
GPU part: protein2datapoints -> motif_both_points_noLSH -> evaluate2 ->
motif_both_points end to end, output files byte-identical to the reference's.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "hsearch_b200", "bin")
REF = os.path.join(ROOT, "oracle", "_ref", "bin")
PROGS = ["motif_both_points", "motif_both_points_noLSH", "protein2datapoints"]
AA = "ARNDCQEGHILKMFPSTWYV"


def _have(dirname, prog):
    return os.access(os.path.join(dirname, prog), os.X_OK)


def _ensure_built():
    if not all(_have(OURS, p) for p in PROGS):
        from hsearch_b200 import build
        build.build()


def run(dirname, prog, args, cwd, env=None):
    """argv[0] is the bare program name for both builds so that the echoed
    command line (motif_both_points.cpp:268-273) compares equal."""
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([prog] + args, executable=os.path.join(dirname, prog), cwd=cwd, env=e,
                       capture_output=True, text=True, timeout=600)
    return p.returncode, p.stdout, p.stderr


needs_ref = pytest.mark.skipif(not all(_have(REF, p) for p in PROGS),
                               reason="oracle/_ref/bin not built (no /root/reference here and no prebuilt copy)")


@needs_ref
@pytest.mark.parametrize("prog", PROGS)
@pytest.mark.parametrize("args", [[], ["-help"], ["-?"], ["-about"], ["-d", "x"], ["-db", "x", "-l", "10"],
                                  ["db", "x"], ["-o", "out", "-l", "7"]])
def test_option_protocol_matches_reference(tmp_path, prog, args):
    _ensure_built()
    want = run(REF, prog, args, tmp_path)
    got = run(OURS, prog, args, tmp_path)
    assert got == want


@pytest.mark.parametrize("prog", PROGS)
def test_missing_required_option_exits_zero(tmp_path, prog):
    """First missing required option -> message on stderr, exit status SUCCESS
    (motif_both_points.cpp:332-335), also without the reference binaries."""
    _ensure_built()
    rc, out, err = run(OURS, prog, ["-l", "10"], tmp_path)
    assert rc == 0
    assert "required argument missing" in err and "-d" in err
    assert out.startswith("[WELCOME TO HSEARCH v1.0]\n[%s -l 10]\n" % prog)


def write_fasta(path, n, plen, seed, with_eq=True):
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        for i in range(n):
            seq = "".join(AA[c] for c in rng.integers(0, 20, size=plen))
            f.write(f">prot{i} some description\n{seq}\n")


@needs_ref
@pytest.mark.parametrize("length,plen", [(10, 30), (25, 40), (8, 8)])
def test_protein2datapoints_matches_reference(tmp_path, length, plen):
    """Proteins no longer than len + 29 yield exactly one window (j = 0), so the
    random stride (rand(), protein2datapoints.cpp:57,70) never matters; E/Q are
    present, so the ProteinDB swap (protein.hpp:59-63) is exercised."""
    _ensure_built()
    write_fasta(tmp_path / "db.fa", 300, plen, seed=5)
    a = run(REF, "protein2datapoints", ["-d", "db.fa", "-l", str(length), "-n", "250", "-o", "ref.pts"], tmp_path)
    b = run(OURS, "protein2datapoints", ["-d", "db.fa", "-l", str(length), "-n", "250", "-o", "our.pts"], tmp_path)
    assert a[0] == b[0] == 0
    ref, our = open(tmp_path / "ref.pts").read(), open(tmp_path / "our.pts").read()
    assert ref == our and ref.count("\n") == 2 * 250

    def strip_time(s):
        return [l for l in s.splitlines() if not l.startswith("It takes")]
    assert strip_time(a[1]) == strip_time(b[1].replace("our.pts", "ref.pts"))


@needs_ref
@pytest.mark.parametrize("length,seed", [(10, 7), (25, 12345)])
def test_protein2datapoints_multi_window_branch(tmp_path, length, seed):
    """Long proteins: several windows per protein at the random stride 30 + rand() % 20 and the
    duplicate-k-mer skip (protein2datapoints.cpp:47-70).  rand() is pinned on both sides: the
    reference binary runs under oracle/_ref/libsrand_shim.so (its srand(time(NULL)) seeds HS_SEED),
    the drop-in honours HS_SEED itself.  Non-amino-acid letters exercise ProteinDB's rand() % 20
    replacement (protein.hpp:58-62) with the same stream."""
    shim = os.path.join(ROOT, "oracle", "_ref", "libsrand_shim.so")
    if not os.path.exists(shim):
        pytest.skip("oracle/_ref/libsrand_shim.so not built")
    _ensure_built()
    rng = np.random.default_rng(seed)
    unit = "".join(AA[c] for c in rng.integers(0, 20, size=200))
    with open(tmp_path / "db.fa", "w") as f:
        for i in range(40):
            plen = int(rng.integers(150, 900))
            seq = "".join(AA[c] for c in rng.integers(0, 20, size=plen))
            if i % 5 == 0:       # repeated material: the same k-mers come back (dedup branch)
                seq = unit + seq[:100] + unit
            if i % 7 == 0:       # letters outside the 20 (B, X, Z ...): replaced with rand() % 20
                seq = seq[:50] + "BXZ" + seq[50:]
            f.write(f">sp|P{i:05d}|name{i} description words\n{seq}\n")
    env = {"HS_SEED": str(seed)}
    a = run(REF, "protein2datapoints", ["-d", "db.fa", "-l", str(length), "-n", "40", "-o", "ref.pts"], tmp_path,
            env=dict(env, LD_PRELOAD=shim))
    b = run(OURS, "protein2datapoints", ["-d", "db.fa", "-l", str(length), "-n", "40", "-o", "our.pts"], tmp_path, env=env)
    assert a[0] == b[0] == 0
    ref, our = open(tmp_path / "ref.pts").read(), open(tmp_path / "our.pts").read()
    assert ref == our
    names = ref.splitlines()[0::2]
    assert len(names) > 200                                   # several windows per protein
    assert any("$0@" not in n for n in names)                 # offsets beyond the first window
    assert names[0].startswith("sp|P00000|name0#0$0@") and names[-1].endswith("*%d" % (len(names) - 1))


def strip_volatile(stdout):
    """Drop the lines that carry timings."""
    keep = []
    for l in stdout.splitlines():
        if l.startswith("ACCURACY:"):
            l = " ".join(l.split()[:2])  # recall only; the second number is seconds
        if l.startswith("It takes") or l.startswith("time "):
            continue
        keep.append(l)
    return keep


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("W", ["50", "20"])
def test_search_pipeline_matches_reference(tmp_path, W):
    """The reference's own workflow, both builds side by side,
    LSH seeded identically through HS_REF_SEED (oracle/fixed_rd.hpp)."""
    _ensure_built()
    length, R = 10, "30"
    write_fasta(tmp_path / "db.fa", 4000, 36, seed=7)
    write_fasta(tmp_path / "q.fa", 60, 12, seed=8)
    # plant near neighbours: queries 0..29 copy DB windows with one substitution
    db = open(tmp_path / "db.fa").read().splitlines()
    q = open(tmp_path / "q.fa").read().splitlines()
    rng = np.random.default_rng(9)
    for i in range(30):
        s = list(db[2 * int(rng.integers(0, 4000)) + 1][:12])
        s[int(rng.integers(0, 10))] = AA[int(rng.integers(0, 20))]
        q[2 * i + 1] = "".join(s)
    open(tmp_path / "q.fa", "w").write("\n".join(q) + "\n")
    env = {"HS_REF_SEED": "12345"}
    for d, tag in ((REF, "ref"), (OURS, "our")):
        assert run(d, "protein2datapoints", ["-d", "db.fa", "-l", str(length), "-n", "4000", "-o", f"{tag}.db"],
                   tmp_path)[0] == 0
        assert run(d, "protein2datapoints", ["-d", "q.fa", "-l", str(length), "-n", "60", "-o", f"{tag}.q"],
                   tmp_path)[0] == 0
    assert open(tmp_path / "ref.db").read() == open(tmp_path / "our.db").read()
    assert open(tmp_path / "ref.q").read() == open(tmp_path / "our.q").read()

    # brute force ground truth
    a = run(REF, "motif_both_points_noLSH", ["-d", "ref.db", "-c", "ref.q", "-l", str(length), "-T", R, "-o", "ref.gt"],
            tmp_path, env)
    b = run(OURS, "motif_both_points_noLSH", ["-d", "our.db", "-c", "our.q", "-l", str(length), "-T", R, "-o", "our.gt"],
            tmp_path, dict(env, HS_NOLSH_NONHITS="1"))
    assert a[0] == 0 and b[0] == 0, b[2]
    ref_gt, our_gt = open(tmp_path / "ref.gt").read(), open(tmp_path / "our.gt").read()
    assert ref_gt == our_gt and ref_gt.count("\n") > 30
    # the dump of every non-hit (motif_both_points_noLSH.cpp:47-49)
    assert open(tmp_path / "ref.gtnotlessthan.txt").read() == open(tmp_path / "our.gtnotlessthan.txt").read()

    # evaluate2 sorts the ground truth (evaluate2.cpp:88-96); both searches read the same sorted file
    ea = run(REF, "evaluate2", ["ref.gt"], tmp_path)
    eb = run(OURS, "evaluate2", ["our.gt"], tmp_path)
    assert ea[0] == eb[0] == 0 and ea[1] == eb[1].replace("our.gt", "ref.gt")
    gts = "ref.gtsort.txt"
    assert os.path.exists(tmp_path / gts)
    assert open(tmp_path / gts).read() == open(tmp_path / "our.gtsort.txt").read()

    a = run(REF, "motif_both_points", ["-d", "ref.db", "-c", "ref.q", "-l", str(length), "-W", W, "-T", R, "-g", gts,
                                       "-o", "ref.hits"], tmp_path, env)
    b = run(OURS, "motif_both_points", ["-d", "our.db", "-c", "our.q", "-l", str(length), "-W", W, "-T", R, "-g", gts,
                                        "-o", "our.hits"], tmp_path, env)
    assert a[0] == 0 and b[0] == 0, b[2]
    ref_hits, our_hits = open(tmp_path / "ref.hits").read(), open(tmp_path / "our.hits").read()
    assert ref_hits == our_hits and ref_hits.count("\n") > 3
    assert open(tmp_path / "ref.hits.accuracy.txt").read() == open(tmp_path / "our.hits.accuracy.txt").read()
    sa, sb = strip_volatile(a[1]), strip_volatile(b[1])
    sb = [l.replace("our.", "ref.") for l in sb]
    assert sa == sb

    # multi-GPU mode of the two search programs (HS_DEVICES: one context per listed device, the kmers
    # sharded in contiguous blocks, lists merged into the reference's order on the host): same files and
    # the same stdout as the reference.  "0,0,0" runs three shards on one GPU.
    ndev = 1
    try:
        import torch
        ndev = max(1, torch.cuda.device_count())
    except Exception:
        pass
    devs = ",".join(str(i % ndev) for i in range(3))
    c = run(OURS, "motif_both_points_noLSH", ["-d", "our.db", "-c", "our.q", "-l", str(length), "-T", R, "-o", "mg.gt"],
            tmp_path, dict(env, HS_DEVICES=devs))
    assert c[0] == 0, c[2]
    assert open(tmp_path / "mg.gt").read() == ref_gt
    c = run(OURS, "motif_both_points", ["-d", "our.db", "-c", "our.q", "-l", str(length), "-W", W, "-T", R, "-g", gts,
                                        "-o", "mg.hits"], tmp_path, dict(env, HS_DEVICES=devs))
    assert c[0] == 0, c[2]
    assert open(tmp_path / "mg.hits").read() == ref_hits
    assert [l.replace("mg.", "ref.").replace("our.", "ref.") for l in strip_volatile(c[1])] == sa   # incl. "table size"


# ---------------------------------------------------------------- hclust2 / hclust3 (CL1)
HCL = ["hclust2", "hclust3"]
needs_ref_hcl = pytest.mark.skipif(not all(_have(REF, p) for p in HCL), reason="oracle/_ref/bin/hclust2 not built")


@needs_ref_hcl
@pytest.mark.parametrize("prog", HCL)
@pytest.mark.parametrize("args", [[], ["-help"], ["-?"], ["-about"], ["-k", "x"], ["-kmers", "x", "-l", "10"],
                                  ["-o", "out", "-l", "7", "-K", "4"]])
def test_hclust_option_protocol_matches_reference(tmp_path, prog, args):
    _ensure_built()
    assert run(OURS, prog, args, tmp_path) == run(REF, prog, args, tmp_path)


def write_kmers(path, n, length, seed, family=6):
    """>name / KMER tokens (hclust2.cpp:232-240): near-duplicate families so that clusters form."""
    rng = np.random.default_rng(seed)
    nfam = max(1, n // family)
    roots = rng.integers(0, 20, size=(nfam, length))
    with open(path, "w") as f:
        for i in range(n):
            row = roots[rng.integers(0, nfam)].copy()
            for _ in range(int(rng.integers(0, 3))):
                row[rng.integers(0, length)] = rng.integers(0, 20)
            f.write(f">kmer{i}\n{''.join(AA[c] for c in row)}\n")


@pytest.mark.gpu
@needs_ref_hcl
@pytest.mark.parametrize("prog,length,K,L,W,R", [("hclust2", 10, 4, 4, "50", "25"), ("hclust3", 10, 4, 6, "30", "20"),
                                                 ("hclust2", 25, 16, 8, "50", "60"), ("hclust2", 8, 2, 3, "20", "30")])
def test_hclust_matches_reference(tmp_path, prog, length, K, L, W, R):
    """Greedy centre clustering (Clustering(), hclust2.cpp:86-151) on the GPU against the
    reference program itself: cluster file byte-identical, stdout equal up to the timings."""
    _ensure_built()
    write_kmers(tmp_path / "kmers.txt", 3000, length, seed=11)
    env = {"HS_REF_SEED": "777"}
    args = ["-k", "kmers.txt", "-l", str(length), "-K", str(K), "-L", str(L), "-W", W, "-T", R]
    a = run(REF, prog, args + ["-o", "ref.clu"], tmp_path, env)
    b = run(OURS, prog, args + ["-o", "our.clu"], tmp_path, env)
    assert a[0] == 0 and b[0] == 0, b[2]
    ref, our = open(tmp_path / "ref.clu").read(), open(tmp_path / "our.clu").read()
    assert ref == our
    nclusters = ref.count("#clusterid:")
    assert 10 < nclusters < 3000  # real merging happened

    def strip(s):
        return [l for l in s.splitlines() if "takes" not in l]
    assert strip(a[1]) == [l.replace("our.clu", "ref.clu") for l in strip(b[1])]


@pytest.mark.skipif(not _have(REF, "evaluate2"), reason="oracle/_ref/bin/evaluate2 not built")
def test_evaluate2_matches_reference(tmp_path):
    """Ground-truth sorter (evaluate2.cpp:75-96): same stdout, same <file>sort.txt."""
    _ensure_built()
    rng = np.random.default_rng(17)
    rows = [(f"motif{int(rng.integers(0, 40))}", f"prot#{i}$0@K*1", float(rng.random() * 60)) for i in range(500)]
    with open(tmp_path / "gt", "w") as f:
        for m, p, d in rows:
            f.write(f"{m} {p} {d:.6g}\n")
    import shutil
    shutil.copy(tmp_path / "gt", tmp_path / "gt2")
    a = run(REF, "evaluate2", ["gt"], tmp_path)
    b = run(OURS, "evaluate2", ["gt2"], tmp_path)
    assert a[0] == b[0] == 0 and a[1] == b[1].replace("gt2", "gt")
    assert open(tmp_path / "gtsort.txt").read() == open(tmp_path / "gt2sort.txt").read()
    # missing input file: empty sorted file, exit 0, like the reference
    a = run(REF, "evaluate2", ["nope"], tmp_path)
    b = run(OURS, "evaluate2", ["nope"], tmp_path)
    assert a[:2] == b[:2] and open(tmp_path / "nopesort.txt").read() == ""


# ---------------------------------------------------------------- orf (E6)
@pytest.mark.gpu
def test_orf_program_matches_oracle(tmp_path, oracle):
    """`orf -q file` (orf/orf_main.cc:8-20): the kept frames of every sequence, named name_j, against the
    oracle's restatement of ORF::orf6 (the reference's own program does not compile: headers missing)."""
    _ensure_built()
    rng = np.random.default_rng(3)
    seqs = ["".join("ACGT"[c] for c in rng.integers(0, 4, size=int(n))) for n in rng.integers(5, 400, size=50)]
    seqs += ["ATG" * 30, "AC", ""]           # a long open frame, a sequence shorter than a codon, an empty one
    with open(tmp_path / "q.fa", "w") as f:
        for i, s in enumerate(seqs):
            f.write(f">read{i} x\n")
            for k in range(0, len(s), 60):    # multi-line FASTA
                f.write(s[k:k + 60] + "\n")
    rc, out, err = run(OURS, "orf", ["-q", "q.fa"], tmp_path)
    assert rc == 0, err
    got = open(tmp_path / "q.fa_translatedAA.fasta").read().splitlines()
    want = []
    for i, s in enumerate(seqs):
        for j, frame in enumerate(oracle.orf6(s)):     # the kept frames (>= 6 residues), in frame order
            want += [f"read{i} x_{j}", frame]
    assert len(want) > 20 and got == want
    assert run(OURS, "orf", [], tmp_path)[0] != 0


# ---------------------------------------------------------------- pcluster (E4 + KL1)
@pytest.mark.gpu
def test_pcluster_program_pregroups(tmp_path, oracle):
    """`pcluster -d fasta -o out` (pcluster.cpp:84-181): FASTA read on the device, 3-mer KLSH pre-clustering,
    the reference's progress lines; the pre-groups must be those of the oracle's restatement of
    Kmer2Integer + KLSH::GetHashValue (pinned to the reference's own in test_sequence.py) over the sequences the host parser (ReadFASTAFile) yields
    under the same srand."""
    import ctypes as C
    from tests.test_sequence import parse_fasta
    _ensure_built()
    rng = np.random.default_rng(8)
    aa = "ARNDCEQGHILKMFPSTWYV"
    lines = []
    for i in range(400):
        lines.append(f">p{i} some description" if i % 4 else f">p{i}")
        n = int(rng.integers(0, 200)) if i % 17 else int(rng.integers(0, 3))     # some shorter than a 3-mer
        seq = "".join(aa[c] for c in rng.integers(0, 20, size=n))
        if i % 11 == 0 and n > 5:
            seq = seq[:3] + "xbz" + seq[3:]                                       # letters the reader replaces
        for k in range(0, len(seq), 60):
            lines.append(seq[k:k + 60])
    text = "\n".join(lines) + "\n"
    (tmp_path / "db.fa").write_text(text)
    rc, out, err = run(OURS, "pcluster", ["-d", "db.fa", "-o", "groups.txt"], tmp_path, env={"HS_SEED": "7"})
    assert rc == 0, err
    names, seqs = parse_fasta(text, seed=7)
    w, t, b = oracle.klsh_generate(512, 16, 0.2)
    want = {}
    for i, s_ in enumerate(seqs):
        if len(s_) < 3:
            continue
        want.setdefault(oracle.klsh_hash(oracle.kmer3_features(s_), w, t, b), []).append(i)
    got = {}
    for line in open(tmp_path / "groups.txt").read().splitlines():
        hv, n, members = line.split("\t")
        got[int(hv)] = [int(x) for x in members.split()]
        assert len(got[int(hv)]) == int(n)
    assert len(want) > 5 and got == want
    assert list(got) == sorted(got)
    e = err.splitlines()
    assert e[0] == "[WELCOME TO PCLUSTER v1.0]" and e[1] == "[pcluster -d db.fa -o groups.txt]"
    assert e[2] == f"[THE TOTAL NUMBER OF PROTEINS IN THE DATABASE IS {len(seqs)}]"
    assert e[3] == f"[NUMBER OF PRE-GROUPS {len(want)}]"
    assert e[4].startswith("[Locality-Sensitive Hashing Pre-Clustering TAKES ")
    assert e[5] == f"[CLUSTERING GROUP 0 of {len(want)}]"
    # the reference's exit protocol: help / missing option on stderr, status 0 (pcluster.cpp:132-143)
    assert run(OURS, "pcluster", [], tmp_path)[0] == 0
    rc, out, err = run(OURS, "pcluster", ["-d", "db.fa"], tmp_path)
    assert rc == 0 and "required argument missing" in err
