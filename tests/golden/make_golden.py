"""Generate the golden fixtures from the REFERENCE ITSELF (oracle/_ref: the
reference's own sources compiled in place by oracle/Makefile).  Run in the build
container, where /root/reference exists:

    python tests/golden/make_golden.py

The fixtures pin both the oracle restatement (tests -m "not gpu") and the CUDA
path (tests -m gpu); /root/reference is never read at test time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.pyoracle import Oracle, Reference  # noqa: E402
from tests.util import planted_families, planted_queries, random_codes  # noqa: E402


def main():
    o, r = Oracle(), Reference()
    out = {}
    # --- H1: projection matrices straight from LSH::LSH -----------------------
    for name, (seed, dim, K, W) in {"proj_a": (12345, 80, 4, 50.0), "proj_b": (99, 200, 8, 7.5),
                                    "proj_c": (2147483647, 24, 3, 1.0)}.items():
        a, b = r.lsh_generate(seed, dim, K, W)
        out[name + "_args"] = np.array([seed, dim, K, W])
        out[name + "_a"], out[name + "_b"] = a, b
    # --- H2-H4: bucket ints and key strings ------------------------------------
    codes = random_codes(600, 10, seed=101)
    out["hash_codes"] = codes
    for tag, print6 in (("p6", True), ("full", False)):
        pts = o.embed(codes, o.coordinates(print6))
        for (K, L, W) in [(4, 4, 50.0), (4, 4, 4.0), (16, 2, 10.0)]:
            b, keys = r.hash_points(pts, K, L, W, 777)
            key = f"hash_{tag}_K{K}_L{L}_W{W:g}"
            out[key + "_buckets"] = b
            out[key + "_keys"] = np.array(keys.tolist(), dtype="U200")
    # --- B1 + V1/V2: Search() as shipped ------------------------------------------
    db = random_codes(6000, 10, seed=102)
    qc = planted_queries(db, 120, seed=103)
    out["search_db"], out["search_q"] = db, qc
    tab = o.coordinates(True)
    for W in (20.0, 50.0):
        hits, printed, ts, _ = r.search(o.embed(db, tab), o.embed(qc, tab), 4, 4, W, 30.0, 12345)
        out[f"search_W{W:g}_hits"] = hits
        out[f"search_W{W:g}_printed"] = printed
        out[f"search_W{W:g}_tsizes"] = ts
    # --- G1: brute force as shipped -------------------------------------------------
    hits, printed, _ = r.bruteforce(o.embed(db, tab), o.embed(qc, tab), 30.0)
    out["brute_hits"], out["brute_printed"] = hits, printed
    # --- U1: UnionFind ------------------------------------------------------------------
    rng = np.random.default_rng(104)
    n = 400
    eu = rng.integers(0, n, size=300).astype(np.uint32)
    ev = rng.integers(0, n, size=300).astype(np.uint32)
    roots = r.union_find_roots(np.arange(n, dtype=np.uint32), eu, ev)
    out["uf_n"], out["uf_eu"], out["uf_ev"], out["uf_roots"] = np.array([n]), eu, ev, roots
    # --- KL1 / E5 ------------------------------------------------------------------------
    fam = planted_families(40, 60, seed=105, family=4, max_sub=6)
    letters = "ARNDCQEGHILKMFPSTWYV"
    seqs = ["".join(letters[c] for c in row) for row in fam]
    feats = np.stack([o.kmer3_features(s) for s in seqs])
    out["klsh_seqs"] = np.array(seqs, dtype="U60")
    out["klsh_hash"] = np.array([r.klsh_hash(f) for f in feats], dtype=np.uint64)
    out["kmer2int"] = np.array([r.kmer2integer(s[:3]) for s in seqs], dtype=np.uint32)
    out["weight_d"] = np.array([0.0, 1e-9, 10.0, 23.999, 24.0, 24.5, 25.0, 26.0, 30.0])
    out["weight_w"] = np.array([r.search_lib.ref_weight(d, 30.0) for d in out["weight_d"]])
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"), len(out), "arrays")


def evaluate_lists(seed=5, n=4000, Q=50, N=2000, R=30.0):
    """Seeded ground-truth / found hit lists for the R1 (recall) fixtures: 60 % of the truth
    found, 30 found pairs absent from the truth, found in (query, first table, db id) order."""
    from oracle.pyoracle import HIT_DTYPE
    rng = np.random.default_rng(seed)
    t = np.zeros(n, dtype=HIT_DTYPE)
    pairs = np.sort(rng.choice(Q * N, size=n, replace=False))
    t["query"], t["db_id"] = pairs // N, pairs % N
    t["dist2"] = rng.uniform(0, R, n) ** 2
    keep = rng.random(n) < 0.6
    f = t[keep].copy()
    f["table_first"] = rng.integers(0, 4, int(keep.sum()))
    e = np.zeros(30, dtype=HIT_DTYPE)
    e["query"], e["db_id"], e["dist2"] = rng.integers(0, Q, 30), N + np.arange(30), 100.0
    e["table_first"] = rng.integers(0, 4, 30)
    f = np.sort(np.concatenate([f, e]), order=["query", "table_first", "db_id"])
    return t, f, Q, R


def main_evaluate():
    """evaulate() of the reference (motif_both_points.cpp:100-165) on seeded lists."""
    r = Reference()
    t, f, Q, R = evaluate_lists()
    recall, rows = r.evaluate(t, f, R)
    tpb, fnb = np.zeros(500, dtype=np.uint64), np.zeros(500, dtype=np.uint64)
    for row in rows:  # "<bin> <ratio> <tp> <fn>" | "<bin> 0 fn <fn>" | "<bin> 1 tp <tp>"  (:151-163)
        b = int(row[0])
        if row[2] == "fn":
            fnb[b] = int(row[3])
        elif row[2] == "tp":
            tpb[b] = int(row[3])
        else:
            tpb[b], fnb[b] = int(row[2]), int(row[3])
    np.savez_compressed(os.path.join(HERE, "evaluate_golden.npz"), truth=t, found=f, Q=Q, R=R, recall=recall,
                        tp_bin=tpb, fn_bin=fnb)
    print("wrote evaluate_golden.npz: recall", recall)


if __name__ == "__main__":
    if "--evaluate-only" not in sys.argv:
        main()
    main_evaluate()
