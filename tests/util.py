"""Shared helpers for the parity tests: seeded synthetic inputs."""
import numpy as np


def random_codes(n, length, seed):
    """i.i.d. uniform over the 20 letters (the reference's only data model,
    BLOSUM-Metric/src/BLOSUM-metric/evaluate.cpp:21-28)."""
    return np.random.default_rng(seed).integers(0, 20, size=(n, length), dtype=np.uint8)


def planted_queries(db_codes, q, seed, frac=0.5, max_sub=2):
    """Random queries, a fraction planted as 1..max_sub-substitution mutants of DB
    fragments so that true neighbours exist."""
    rng = np.random.default_rng(seed)
    n, length = db_codes.shape
    qc = rng.integers(0, 20, size=(q, length), dtype=np.uint8)
    nplant = int(q * frac)
    src = rng.integers(0, n, size=nplant)
    for i in range(nplant):
        row = db_codes[src[i]].copy()
        for _ in range(int(rng.integers(0, max_sub + 1))):
            row[rng.integers(0, length)] = rng.integers(0, 20)
        qc[i] = row
    return qc


def planted_families(n, length, seed, family=8, max_sub=2):
    """DB made of near-duplicate families (cluster tests)."""
    rng = np.random.default_rng(seed)
    nfam = max(1, n // family)
    roots = rng.integers(0, 20, size=(nfam, length), dtype=np.uint8)
    out = np.zeros((n, length), dtype=np.uint8)
    for i in range(n):
        row = roots[i % nfam].copy()
        for _ in range(int(rng.integers(0, max_sub + 1))):
            row[rng.integers(0, length)] = rng.integers(0, 20)
        out[i] = row
    return out


def hits_as_tuples(h, with_table=True):
    if with_table:
        return list(zip(h["query"].tolist(), h["table_first"].tolist(), h["db_id"].tolist(), h["dist2"].tolist()))
    return list(zip(h["query"].tolist(), h["db_id"].tolist(), h["dist2"].tolist()))
