#!/usr/bin/env python
"""Times the search stages of bench C2 (10 k queries x N fragments) for the library selected with
HS_LIBRARY; prints one JSON line (filter / exact / hit-sort ms, survivors, hits, hit checksum)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hsearch_b200 as hb  # noqa: E402

N = int(os.environ.get("FT_N", 100_000_000))
Q, length = int(os.environ.get("FT_Q", 10_000)), int(os.environ.get("FT_LEN", 10))
reps = int(os.environ.get("FT_REPS", 3))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1000)
codes = torch.randint(0, 20, (N, length), dtype=torch.uint8, device=dev, generator=g)
table = torch.tensor(hb.coordinates(hb.HS_TABLE_PRINT6), dtype=torch.float64, device=dev)
qcodes = torch.randint(0, 20, (Q, length), dtype=torch.uint8, device=dev, generator=g)
qpts = table[qcodes.long()].reshape(Q, 8 * length).contiguous()
h = hb.HSearch(length, 4, 4, 50.0, float(os.environ.get("FT_R", 30.0)), flags=hb.HS_FLAG_SORT_HITS)
h.seed_projection(12345)
h.load_fragments_dev(codes.data_ptr(), N)
h.build_index()
sb = h.stats().as_dict()
nh = h.search_points_dev(qpts.data_ptr(), Q, 0, 0)
cap = int(nh * 1.05) + 1024
buf = torch.empty(cap * 24, dtype=torch.uint8, device=dev)
out = []
for _ in range(reps):
    h.search_points_dev(qpts.data_ptr(), Q, buf.data_ptr(), cap)
    out.append(h.stats().as_dict())
s = min(out, key=lambda d: d["ms_filter_tc"])
print(json.dumps({"lib": os.path.basename(os.environ.get("HS_LIBRARY", "default")), "N": N,
                  "filter_tc_ms": round(s["ms_filter_tc"], 3), "filter_ms": round(s["ms_filter"], 3),
                  "exact_ms": round(s["ms_exact"], 3), "hitsort_ms": round(s["ms_hitsort"], 3),
                  "search_total_ms": round(s["ms_total"], 3), "hash_ms": round(sb["ms_hash"], 3),
                  "sort_ms": round(sb["ms_sort"], 3), "group_ms": round(sb["ms_group"], 3), "permute_ms": round(sb["ms_permute"], 3),
                  "candidates": s["n_candidates"], "survivors": s["n_survivors"], "hits": s["n_hits"],
                  "checksum": "%016x" % h.hits_checksum_dev(buf.data_ptr(), min(nh, cap))}))
