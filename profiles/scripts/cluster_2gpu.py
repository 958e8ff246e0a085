"""Sharded near-pair clustering on N GPUs (one process per GPU, torchrun): every rank owns a
block of a synthetic family DB; labels must equal a single-GPU hs_cluster of the whole DB.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 profiles/scripts/cluster_2gpu.py [n_total]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hsearch_b200 as hb  # noqa: E402
from hsearch_b200 import dist as hdist  # noqa: E402
from tests.util import planted_families  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
length, K, L, W, R = 10, 4, 4, 20.0, 25.0
codes = planted_families(n_total, length, seed=5) if n_total <= 200_000 else None
if codes is None:  # large: vectorised generator (families of 8 with <= 2 substitutions)
    rng = np.random.default_rng(5)
    roots = rng.integers(0, 20, size=(n_total // 8 + 1, length), dtype=np.uint8)
    codes = roots[np.arange(n_total) % len(roots)].copy()
    for _ in range(2):
        m = rng.random(n_total) < 0.5
        pos = rng.integers(0, length, size=n_total)
        val = rng.integers(0, 20, size=n_total, dtype=np.uint8)
        codes[np.nonzero(m)[0], pos[m]] = val[m]
lo, hi = hdist.shard_range(n_total, rank, world)
h = hb.HSearch(length, K, L, W, R, predicate=hb.HS_PRED_SQRT_LE_R, flags=0, device=local)
a, b = h.seed_projection(777)
h.load_fragments(codes[lo:hi], id_base=lo)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
h.hash()
key_fn, edges_fn, union_fn = hdist.gpu_cluster_callbacks(h, a, b, codes[lo:hi])
labels = hdist.cluster_sharded(codes[lo:hi], lo, n_total, L, key_fn, edges_fn, union_fn, device=f"cuda:{local}")
torch.cuda.synchronize()
dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    g = hb.HSearch(length, K, L, W, R, predicate=hb.HS_PRED_SQRT_LE_R, flags=0, device=local)
    g.set_projection(a, b)
    g.load_fragments(codes)
    g.build_index()
    t1 = time.perf_counter()
    want = g.cluster()
    t_single = time.perf_counter() - t1
    ok = np.array_equal(labels, want[lo:hi])
    print(f"cluster_sharded world={world} n_total={n_total} clusters={len(np.unique(want))} "
          f"match_single_gpu={ok} sharded_s={dt:.3f} single_gpu_cluster_s={t_single:.3f}", flush=True)
    assert ok
dist.destroy_process_group()
