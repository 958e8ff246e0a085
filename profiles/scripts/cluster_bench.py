"""configs[3] at reduced size: near-pair clustering (hs_cluster) of planted families, K = 4, L = 8,
W = 50, R = 25, tensor-filter self-join vs the scalar filter."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hsearch_b200 as hb  # noqa: E402


def families(n, length, seed):
    rng = np.random.default_rng(seed)
    roots = rng.integers(0, 20, size=(n // 10 + 1, length), dtype=np.uint8)
    codes = roots[np.arange(n) % len(roots)].copy()
    for _ in range(2):
        m = rng.random(n) < 0.5
        pos = rng.integers(0, length, size=n)
        val = rng.integers(0, 20, size=n, dtype=np.uint8)
        codes[np.nonzero(m)[0], pos[m]] = val[m]
    return codes


for n in [int(x) for x in sys.argv[1:]] or [1_000_000]:
    codes = families(n, 10, 5)
    for name, flags in (("tensor", 0), ("scalar", hb.HS_FLAG_SCALAR_FILTER)):
        if name == "scalar" and n > 2_000_000:
            continue
        h = hb.HSearch(10, 4, 8, 50.0, 25.0, predicate=hb.HS_PRED_SQRT_LE_R, flags=flags)
        h.seed_projection(12345)
        h.load_fragments(codes)
        h.build_index()
        t0 = time.perf_counter()
        lab = h.cluster()
        wall = time.perf_counter() - t0
        s = h.stats().as_dict()
        print(json.dumps({"config": "C4 cluster", "n": n, "filter": name, "pairs": s["n_candidates"],
                          "pairs_tensor": s["n_candidates_tc"], "survivors": s["n_survivors"], "edges": s["n_edges"],
                          "clusters": int(len(np.unique(lab))), "device_ms": round(s["ms_total"], 1),
                          "wall_ms": round(wall * 1e3, 1), "pairs_per_s": round(s["n_candidates"] / (s["ms_total"] * 1e-3))}),
              flush=True)
        h.close()
