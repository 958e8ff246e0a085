"""Timings of the BASELINE.json configurations other than the bench line (configs[0], [4] and
the K x L sweep of configs[1]); parity for all of them is in tests/, this only records speed.
    python profiles/scripts/configs_bench.py > gpurun_out/configs.jsonl
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hsearch_b200 as hb  # noqa: E402
from tests.util import planted_queries, random_codes  # noqa: E402


def out(**kw):
    print(json.dumps(kw), flush=True)


def search_config(name, n, q, length, K, L, W, R):
    codes = random_codes(n, length, seed=1)
    qcodes = planted_queries(codes[:min(n, 200000)], q, seed=2)
    h = hb.HSearch(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS)
    h.seed_projection(12345)
    h.load_fragments(codes)
    h.build_index()
    hits = h.search_codes(qcodes, cap=1 << 22)           # warm-up + sizes the buffers
    t0 = time.perf_counter()
    h.hash()
    s_hash = h.stats().as_dict()
    h.build_index()
    s_build = h.stats().as_dict()
    hits = h.search_codes(qcodes, cap=len(hits) + 1024)
    s_search = h.stats().as_dict()
    wall = time.perf_counter() - t0
    # device time of the stages (the numpy hit buffer is pageable, so the copy-out inside
    # ms_total runs at a few GB/s and says nothing about the path)
    srch = sum(s_search[k] for k in ("ms_qhash", "ms_probe", "ms_host", "ms_filter", "ms_exact", "ms_hitsort"))
    dev_ms = s_hash["ms_hash"] + s_build["ms_total"] + srch
    out(config=name, n_db=n, n_query=q, len=length, K=K, L=L, W=W, R=R, rank_path=bool(s_build["rank_path"]),
        key_words=s_build["key_words"], candidates=s_search["n_candidates"], survivors=s_search["n_survivors"],
        hits=int(len(hits)), device_ms=round(dev_ms, 3), wall_ms=round(wall * 1e3, 2),
        db_fragments_per_s=round(n / (dev_ms * 1e-3)), hash_ms=round(s_hash["ms_hash"], 3),
        build_ms=round(s_build["ms_total"], 3), sort_passes=s_build["sort_passes"], search_ms=round(srch, 3),
        filter_ms=round(s_search["ms_filter"], 3), exact_ms=round(s_search["ms_exact"], 3),
        hitsort_ms=round(s_search["ms_hitsort"], 3), qhash_ms=round(s_search["ms_qhash"], 3),
        probe_ms=round(s_search["ms_probe"], 3), host_ms=round(s_search["ms_host"], 3),
        sort_ms=round(s_build["ms_sort"], 3), group_ms=round(s_build["ms_group"], 3),
        permute_ms=round(s_build["ms_permute"], 3), hash_sort_fallbacks=s_build["hash_sort_fallbacks"])
    h.close()


def allpairs_config(n, length, metric, R):
    codes = random_codes(n, length, seed=3)
    h = hb.HSearch(length, 4, 4, 50.0, R, metric=metric, predicate=hb.HS_PRED_SQRT_LE_R, flags=0)
    h.load_fragments(codes)
    hits = h.bruteforce_codes(None, cap=1 << 22)         # warm-up
    t0 = time.perf_counter()
    hits = h.bruteforce_codes(None, cap=max(len(hits), 1) + 1024)
    wall = time.perf_counter() - t0
    s = h.stats().as_dict()
    pairs = n * (n - 1) // 2
    out(config="C5 all-pairs", n=n, len=length, metric="blosum_int" if metric == hb.HS_METRIC_BLOSUM_INT else "euclid_fp64",
        R=R, pairs=pairs, survivors=s["n_survivors"], hits=int(len(hits)), device_ms=round(s["ms_total"], 2),
        wall_ms=round(wall * 1e3, 2), pairs_per_s=round(pairs / (s["ms_total"] * 1e-3)),
        filter_ms=round(s["ms_filter"], 2), exact_ms=round(s["ms_exact"], 2))
    h.close()


if __name__ == "__main__":
    if len(sys.argv) > 1:   # only the named sweep points: K,L,W ...
        for t in sys.argv[1:]:
            if t == "allpairs":
                n = int(os.environ.get("HS_C5_N", "1000000"))
                for length in (8, 10, 12, 16, 20, 25, 30):
                    allpairs_config(n, length, hb.HS_METRIC_BLOSUM_INT, float(3 * length))
                continue
            K, L, W = t.split(",")
            search_config("C2 sweep", 10_000_000, 10_000, 10, int(K), int(L), float(W), 30.0)
        sys.exit(0)
    search_config("C1 (configs[0])", 1_000_000, 1000, 10, 4, 4, 50.0, 30.0)
    search_config("C1 W=20", 1_000_000, 1000, 10, 4, 4, 20.0, 30.0)
    for K, L, W in [(2, 1, 50.0), (2, 4, 20.0), (4, 1, 50.0), (4, 4, 50.0), (4, 16, 50.0), (4, 4, 20.0), (4, 4, 10.0),
                    (8, 4, 50.0), (8, 16, 20.0), (16, 4, 50.0), (16, 32, 20.0)]:
        search_config("C2 sweep", 10_000_000, 10_000, 10, K, L, W, 30.0)
    n = int(os.environ.get("HS_C5_N", "1000000"))
    for length in (8, 10, 12, 16, 20, 25, 30):
        allpairs_config(n, length, hb.HS_METRIC_BLOSUM_INT, float(3 * length))
    for length, R in ((10, 24.0), (25, 40.0)):
        allpairs_config(n, length, hb.HS_METRIC_EUCLID_FP64, R)
