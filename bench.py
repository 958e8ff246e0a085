#!/usr/bin/env python
"""bench.py -- the HSEARCH hot path on B200: DB fragments hashed + bucketed +
verified per second (BASELINE.json `metric`), on BASELINE.json configs[1]
(10k queries vs 100M synthetic fragments per GPU, len 10, K = L = 4).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU code

One "step" is one pass of the whole path over one batch: hash every fragment
(K1), radix-sort + group + bucket-order the L tables (K2), hash / probe / filter
/ exactly verify the Q queries and put the hits in the reference's order (K3).
`value` times that with the DB codes and queries resident in HBM; `e2e` times
the same through the C ABI with host buffers (H2D of the codes and queries and
D2H of the hits inside the timed region).

Only the `cpu_baseline` leg and `--impl reference` touch oracle/ (the checker);
the product path is libhsearch_b200.so and fails loudly without a GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "db_fragments_hashed_bucketed_verified_per_s"
UNIT = "fragments/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--n-db", type=int, default=100_000_000, help="DB fragments per GPU")
    ap.add_argument("--n-query", type=int, default=10_000)
    ap.add_argument("--len", type=int, default=10)
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--L", type=int, default=4)
    ap.add_argument("--W", type=float, default=50.0)
    ap.add_argument("--R", type=float, default=30.0)
    ap.add_argument("--planted", type=float, default=0.1, help="fraction of queries planted as DB mutants")
    ap.add_argument("--cpu-sample", type=int, default=20_000, help="DB fragments per host core in the CPU leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-contexts", type=int, default=2, help="contexts alternating batches in the end-to-end leg (1 or 2)")
    ap.add_argument("--no-recall", action="store_true", help="skip the full-size brute-force recall measurement")
    ap.add_argument("--audit", action="store_true", help="FP64 audit inside every timed hs_hash too (the bench always audits "
                    "the keys once after the timed region)")
    ap.add_argument("--no-subset-check", action="store_true", help="skip the exact CPU comparison of a slice of the result")
    ap.add_argument("--subset-db", type=int, default=1_000_000)
    ap.add_argument("--subset-q", type=int, default=1000)
    ap.add_argument("--scalar-filter", action="store_true", help="A/B: keep all candidates on the scalar filter")
    ap.add_argument("--workload", choices=["search", "cluster"], default="search",
                    help="search = the headline (configs[1]/[2]); cluster = configs[3], near-pair union-find of "
                         "--n-db fragments in total (K=4 L=8 W=50 R=25 unless given), pair work split over the GPUs")
    return ap.parse_args()


def workload_name(a):
    return (f"LSH search: {a.n_query} queries vs {a.n_db} synthetic fragments per GPU, len {a.len}, "
            f"K={a.K} L={a.L} W={a.W:g} R={a.R:g}")


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons every 50 ms.  Started before the warm-up steps (nvidia-smi
    needs a few hundred ms to come up, longer than a short timed region); stop(t0, t1) keeps the
    samples whose timestamps fall inside the timed region [t0, t1] (time.time() seconds) and falls
    back to all samples under load (warm-up + timed steps, the same workload) when none does."""
    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self, t0=None, t1=None):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)  # let the sample covering the end of the region be written
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        import datetime
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[2]), float(c[3]), [n for n, v in zip(names, c[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if t0 is not None and t0 - 0.05 <= r[0] <= t1 + 0.05]
        window = "timed region"
        if not inside:
            inside, window = rows, "warm-up + timed steps (no sample fell inside the timed region)"
        sm = [r[1] for r in inside]
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[2] for r in inside) if inside else None,
                "samples": len(sm), "window": window, "reasons": reasons}


# ------------------------------------------------------------------ CPU legs
def _cpu_worker(args):
    """One host core: the reference's own Search() (oracle/_ref) -- or the oracle
    port when oracle/_ref is absent -- on a disjoint DB sample."""
    wid, n, q, length, K, L, W, R, planted = args
    from oracle.pyoracle import Oracle, Reference
    from tests.util import planted_queries, random_codes
    o = Oracle()
    tab = o.coordinates(True)
    db = random_codes(n, length, seed=1000 + wid)
    qc = planted_queries(random_codes(n, length, seed=1000), q, seed=2, frac=planted)
    dbp, qp = o.embed(db, tab), o.embed(qc, tab)
    if Reference.available():
        r = Reference()
        hits, _, _, sec = r.search(dbp, qp, K, L, W, R, 12345, cap=1 << 22)
        return "reference", sec, len(hits)
    a, b = o.lsh_tables(12345, 8 * length, K, L, W)
    t0 = time.perf_counter()
    hits, _, _ = o.search(dbp, qp, a, b, W, R, pred=0, cap=1 << 22)
    return "port", time.perf_counter() - t0, len(hits)


def cpu_leg(a, n_sample, cores=None):
    """Runs one process per host core on disjoint DB samples with the full query
    set; returns aggregate fragments/s.  Text parsing excluded on both sides."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    work = [(w, n_sample, a.n_query, a.len, a.K, a.L, a.W, a.R, a.planted) for w in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_cpu_worker, work)
    wall = time.perf_counter() - t0
    kind = res[0][0]
    slowest = max(r[1] for r in res)  # all cores run concurrently; the job ends with the slowest
    value = cores * n_sample / slowest
    return {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": (f"{cores} processes x {n_sample} DB fragments each vs all {a.n_query} queries, "
                       f"Search() seconds (clock(), index build + query loop), slowest core {slowest:.2f} s, "
                       f"wall {wall:.1f} s incl. input generation"),
            "seconds": slowest}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = min(a.cpu_sample, 10_000)
    vals, last = [], None
    for i in range(a.warmup + a.steps):
        last = cpu_leg(a, n_sample)
        if i >= a.warmup:
            vals.append(last)
    secs = [v["seconds"] for v in vals]
    value = last["cores"] * n_sample / (sum(secs) / len(secs))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(a), "sample": last["sample"],
                      "note": ("the reference holds 8*len doubles per fragment (640 B) and cannot hold the workload's database: "
                               "every host core runs the reference's own Search() on a %d-fragment shard against all queries; "
                               "its per-query memset of 4 bytes per fragment (motif_both_points.cpp:225) is nearly free at this "
                               "size, so the ratio against this arm is a conservative extrapolation" % n_sample)},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"],
                            "sample": last["sample"]},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------ native
def run_native(a):
    import torch
    import torch.distributed as dist

    import hsearch_b200 as hb
    from hsearch_b200 import dist as hdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    length, dim, Q, N = a.len, 8 * a.len, a.n_query, a.n_db
    flags = hb.HS_FLAG_SORT_HITS | (hb.HS_FLAG_HASH_AUDIT if a.audit else 0) | \
        (hb.HS_FLAG_SCALAR_FILTER if a.scalar_filter else 0)
    h = hb.HSearch(length, a.K, a.L, a.W, a.R, table_variant=hb.HS_TABLE_PRINT6, flags=flags, device=local)
    h.seed_projection(12345)
    stream = torch.cuda.ExternalStream(h.stream_ptr(), device=dev)

    # synthetic inputs: i.i.d. uniform over the 20 letters, per-rank seed
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    codes = torch.randint(0, 20, (N, length), dtype=torch.uint8, device=dev, generator=g)
    table = torch.tensor(hb.coordinates(hb.HS_TABLE_PRINT6), dtype=torch.float64, device=dev)
    qcodes = torch.randint(0, 20, (Q, length), dtype=torch.uint8, device=dev, generator=g)
    nplant = int(Q * a.planted)
    if nplant:
        src = torch.randint(0, N, (nplant,), device=dev, generator=g)
        mut = codes[src].clone()
        for _ in range(2):
            pos = torch.randint(0, length, (nplant,), device=dev, generator=g)
            val = torch.randint(0, 20, (nplant,), dtype=torch.uint8, device=dev, generator=g)
            keep = torch.rand(nplant, device=dev, generator=g) < 0.5
            cur = mut[torch.arange(nplant, device=dev), pos]
            mut[torch.arange(nplant, device=dev), pos] = torch.where(keep, cur, val)
        qcodes[:nplant] = mut
    qpts = table[qcodes.long()].reshape(Q, dim).contiguous()
    if world > 1:
        hdist.broadcast_queries(qpts, 0)
    torch.cuda.synchronize()

    h.load_fragments_dev(codes.data_ptr(), N, id_base=rank * N)

    # size the hit buffer with one untimed pass
    h.build_index()
    nh0 = h.search_points_dev(qpts.data_ptr(), Q, 0, 0)
    if world > 1:   # the same capacity on every rank
        mx = torch.tensor([nh0], dtype=torch.int64, device=dev)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        cap = int(mx.item() * 1.05) + 1024
    else:
        cap = int(nh0 * 1.05) + 1024
    # N > 1: the contexts join an NCCL communicator inside the library (hs_comm_init; the unique id
    # travels over torch.distributed, the launcher's plumbing); every search then broadcasts rank 0's
    # queries and merges all ranks' hits into rank 0's memory in the reference's order (comm.cu: counts
    # exchanged with NCCL, hits written by their producers straight to their final positions over
    # NVLink), asynchronously, beside the hash / index build of the next batch.
    import ctypes as C
    from hsearch_b200 import capi
    lib = capi.load()
    nslot = 1
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = np.zeros(128, dtype=np.uint8)
            capi.check(lib.hs_comm_unique_id(raw.ctypes.data_as(C.c_void_p)))
            uid = torch.from_numpy(raw)
        uid = uid.to(dev)
        dist.broadcast(uid, 0)
        raw = uid.cpu().numpy()
        capi.check(lib.hs_comm_init(h.ctx, raw.ctypes.data_as(C.c_void_p), rank, world))
        tot = torch.tensor([nh0], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        capi.check(lib.hs_comm_reserve(h.ctx, int(tot.item() * 1.05) + 4096))
    hits_bufs = [torch.empty(cap * 24, dtype=torch.uint8, device=dev)]
    step_no = [0]
    merged = {"ptr": 0, "total": 0}

    acc = {}

    def add_stats(prefix_keys):
        s = h.stats().as_dict()
        for k in prefix_keys:
            acc[k] = acc.get(k, 0.0) + s[k]
        acc["kernel_launches"] = acc.get("kernel_launches", 0) + s["kernel_launches"]
        return s

    trace = {"hash": 0.0, "build": 0.0, "search": 0.0, "search_dev_ms": 0.0, "sort": 0.0, "group": 0.0, "permute": 0.0,
             "filter": 0.0, "exact": 0.0, "hitsort": 0.0} if os.environ.get("HS_BENCH_TRACE") else None

    def step(collect):
        if trace is not None and collect:
            t_a = time.perf_counter()
            h.hash()
            t_b = time.perf_counter()
            h.build_index()
            t_c = time.perf_counter()
            sb_ = h.stats().as_dict()
            for k_ in ("sort", "group", "permute"):
                trace[k_] += sb_["ms_" + k_]
            t_c = time.perf_counter()
            n = h.search_points_dev(qpts.data_ptr(), Q, hits_bufs[0].data_ptr(), cap)
            t_d = time.perf_counter()
            ss_ = h.stats().as_dict()
            for k_ in ("filter", "exact", "hitsort"):
                trace[k_] += ss_["ms_" + k_]
            trace["hash"] += 1e3 * (t_b - t_a)
            trace["build"] += 1e3 * (t_c - t_b)
            trace["search"] += 1e3 * (t_d - t_c)
            trace["search_dev_ms"] += h.stats().as_dict()["ms_total"]
            step_no[0] += 1
            s0 = h.stats().as_dict()
            return n, n, s0, s0, s0
        h.hash()
        s_hash = add_stats(["ms_hash"]) if collect else h.stats().as_dict()
        h.build_index()
        s_build = add_stats(["ms_sort", "ms_group", "ms_permute", "ms_sort_upsweep", "ms_sort_scan",
                             "ms_sort_downsweep"]) if collect else h.stats().as_dict()
        step_no[0] += 1
        buf = hits_bufs[0]
        n = h.search_points_dev(qpts.data_ptr(), Q, buf.data_ptr(), cap)   # N > 1: also starts the merge on rank 0
        s_search = add_stats(["ms_qhash", "ms_probe", "ms_host", "ms_filter", "ms_filter_tc", "ms_exact", "ms_hitsort"]) if collect \
            else h.stats().as_dict()
        return n, n, s_search, s_build, s_hash

    def drain():
        """N > 1: completes the merge of the latest batch (earlier ones completed before it, in stream order)."""
        if world > 1:
            p, t = C.c_void_p(), C.c_uint64(0)
            capi.check(lib.hs_comm_result(h.ctx, C.byref(p), C.byref(t)))
            merged["ptr"], merged["total"] = p.value or 0, int(t.value)

    with torch.cuda.stream(stream):
        sampler = ClockSampler(local)
        sampler.start()
        for _ in range(a.warmup):
            step(False)
        drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        wall0 = time.time()
        e0.record(stream)
        for _ in range(a.steps):
            nh, nh_total, s_search, s_build, s_hash = step(True)
        drain()                            # every batch's hits are on rank 0 before the clock stops
        if world > 1:
            nh_total = merged["total"]
        e1.record(stream)
        torch.cuda.synchronize()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        wall1 = time.time()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop(wall0, wall1)
    dev_ms = e0.elapsed_time(e1)
    if trace is not None:
        print("[trace] rank %d per step: hash %.2f build %.2f (sort %.2f group %.2f permute %.2f) search %.2f (filter %.2f exact %.2f "
              "hitsort %.2f; device total %.2f) ms; step %.2f" %
              (rank, trace["hash"] / a.steps, trace["build"] / a.steps, trace["sort"] / a.steps, trace["group"] / a.steps,
               trace["permute"] / a.steps, trace["search"] / a.steps, trace["filter"] / a.steps, trace["exact"] / a.steps,
               trace["hitsort"] / a.steps, trace["search_dev_ms"] / a.steps, dev_ms / a.steps), file=sys.stderr, flush=True)
        if world > 1:
            dist.barrier()
        raise SystemExit(0)
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = t.tolist()
    ms_per_step = dev_ms / a.steps
    value = N * world / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with host buffers ------------------------
    # Every step: H2D of the step's codes and queries from pinned host memory, index build, search,
    # D2H of the step's hits (compact 12-byte layout, hs_search_points_compact) -- all inside the timed
    # region.  Two contexts on the same GPU alternate batches from two host threads (double
    # buffering): batch i's hits leave over PCIe while batch i+1's codes arrive on the other
    # direction and its kernels run.  `sequential` is the same call sequence on one context.
    e2e = None
    if not a.no_e2e:
        import threading
        host_codes = torch.empty((N, length), dtype=torch.uint8).pin_memory()
        host_codes.copy_(codes)
        host_q = torch.empty((Q, dim), dtype=torch.float64).pin_memory()
        host_q.copy_(qpts)
        torch.cuda.synchronize()
        nctx = 1 if a.e2e_contexts < 2 else 2
        # (a context that joined the communicator gathers its hits on rank 0 instead: N > 1 uses fresh ones)
        extra = []
        for _ in range(nctx if world > 1 else nctx - 1):
            hx = hb.HSearch(length, a.K, a.L, a.W, a.R, table_variant=hb.HS_TABLE_PRINT6, flags=flags, device=local)
            hx.seed_projection(12345)
            extra.append(hx)
        lanes = []
        for hh in (extra if world > 1 else [h] + extra):
            off = torch.empty(Q + 1, dtype=torch.int64).pin_memory()
            idt = torch.empty(cap, dtype=torch.int32).pin_memory()
            d2 = torch.empty(cap, dtype=torch.float64).pin_memory()
            ch = capi.CompactHits(C.cast(off.data_ptr(), C.POINTER(C.c_uint64)), C.cast(idt.data_ptr(), C.POINTER(C.c_uint32)),
                                  C.cast(d2.data_ptr(), C.POINTER(C.c_double)), cap, 0)
            lanes.append({"h": hh, "off": off, "idt": idt, "d2": d2, "ch": ch, "n": C.c_uint64(0),
                          "stream": torch.cuda.ExternalStream(hh.stream_ptr(), device=dev), "err": None})

        def e2e_step(ln):
            hh = ln["h"]
            capi.check(lib.hs_load_fragments(hh.ctx, C.cast(host_codes.data_ptr(), C.POINTER(C.c_uint8)), N, rank * N))
            capi.check(lib.hs_build_index(hh.ctx))
            capi.check(lib.hs_search_points_compact(hh.ctx, C.cast(host_q.data_ptr(), C.POINTER(C.c_double)), Q,
                                                    C.byref(ln["ch"]), C.byref(ln["n"])))

        def lane_run(ln, nsteps):
            try:
                torch.cuda.set_device(local)
                for _ in range(nsteps):
                    e2e_step(ln)
            except Exception as e:   # reported after the join
                ln["err"] = e

        def timed(nsteps_per_lane):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            f0 = torch.cuda.Event(enable_timing=True)
            f0.record(lanes[0]["stream"])
            th = [threading.Thread(target=lane_run, args=(ln, k)) for ln, k in zip(lanes, nsteps_per_lane)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            ends = []
            for ln in lanes:
                f1 = torch.cuda.Event(enable_timing=True)
                f1.record(ln["stream"])
                ends.append(f1)
            torch.cuda.synchronize()
            for ln in lanes:
                if ln["err"] is not None:
                    raise ln["err"]
            te = torch.tensor([max(f0.elapsed_time(f1) for f1 in ends)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return te.item()

        timed([1] * len(lanes))                                  # warm the pinned path of every context
        split = [a.steps - a.steps // 2, a.steps // 2] if nctx == 2 else [a.steps]
        e_ms = timed(split) / a.steps
        seq_ms = timed([a.steps] + [0] * (len(lanes) - 1)) / a.steps   # one context, no overlap across batches
        s_e = lanes[0]["h"].stats().as_dict()  # the last end-to-end search call of context 0
        nh_e = int(lanes[0]["n"].value)
        # the compact result of the last batch expands to the records of the device-resident run
        e2e_same = None
        try:
            if rank != 0:
                raise StopIteration
            exp = np.zeros(max(nh_e, 1), dtype=capi.HIT_DTYPE)
            capi.check(lib.hs_expand_hits(C.byref(lanes[0]["ch"]), Q, rank * N, exp.ctypes.data))
            dev_hits = hits_bufs[(step_no[0] - 1) % nslot][:min(int(nh), cap) * 24].cpu().numpy().view(capi.HIT_DTYPE)
            e2e_same = bool(nh_e == int(nh) and np.array_equal(exp[:nh_e], dev_hits))
            del exp, dev_hits
        except StopIteration:
            pass
        except Exception as e:
            e2e_same = repr(e)
        d2h = int(nh_e * 12 + (Q + 1) * 8)
        e2e = {"value": N * world / (e_ms * 1e-3), "unit": UNIT, "ms_per_step": e_ms,
               "sequential_ms_per_step": seq_ms, "contexts": nctx,
               "search_stages_ms": {k[3:]: round(s_e[k], 3) for k in ("ms_qhash", "ms_probe", "ms_host", "ms_filter",
                                                                        "ms_exact", "ms_hitsort", "ms_total")},
               "h2d_bytes_per_step": int(N * length + Q * dim * 8), "d2h_bytes_per_step": d2h,
               "hits_equal_device_run_after_expansion": e2e_same,
               "api": ("hs_load_fragments + hs_build_index + hs_search_points_compact (pinned host buffers; per-query CSR, "
                       "12 bytes per hit); %d context(s) on the GPU alternate batches from host threads so that one "
                       "batch's D2H overlaps the next batch's H2D and kernels; every copy is inside the timed region"
                       % nctx)}
        for ln in lanes:
            ln["ch"] = None
        del host_codes, lanes
        for hx in extra:
            hx.close()

    # ---- N > 1: the merged list on rank 0 against the ranks' own lists ----------------------------
    # order-independent 64-bit checksum of (query, first table, db id, dist2 bits): the sum of the ranks'
    # checksums must equal the checksum of the list merged on rank 0, which must be in the reference's order
    multi = None
    if world > 1:
        cs = torch.tensor([h.hits_checksum_dev(hits_bufs[0].data_ptr(), min(int(nh), cap)) - (1 << 63)], dtype=torch.int64, device=dev)
        parts = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(parts, cs)
        cnts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(cnts, torch.tensor([int(nh)], dtype=torch.int64, device=dev))
        if rank == 0:
            want = sum(int(p.item()) + (1 << 63) for p in parts) % (1 << 64)
            got = h.hits_checksum_dev(merged["ptr"], merged["total"])

            class _Raw:
                def __init__(self, ptr, nbytes):
                    self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
            mv = torch.as_tensor(_Raw(merged["ptr"], merged["total"] * 24), device=dev)
            ib = max(1, (world * N - 1).bit_length())
            mkey = (mv.view(torch.int32)[0::6].to(torch.int64) << (ib + 6)) | (mv.view(torch.int32)[1::6].to(torch.int64) << ib) | \
                mv.view(torch.int64)[1::3]
            m_order = bool((mkey[1:] > mkey[:-1]).all().item()) if merged["total"] > 1 else True
            del mkey, mv
            multi = {"ranks": world, "hits_per_rank": [int(c.item()) for c in cnts], "merged_hits_on_rank0": merged["total"],
                     "sum_of_rank_checksums": "%016x" % want, "merged_list_checksum": "%016x" % got,
                     "checksums_equal": bool(want == got and merged["total"] == sum(int(c.item()) for c in cnts)),
                     "merged_list_in_reference_order": m_order}

    # ---- FP64 audit of the keys of the timed run (every projection of every fragment) -----------
    h.hash()
    flips = torch.tensor([h.hash_audit()], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(flips)
    residual_flips = int(flips.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    KW = s_search["key_words"]
    steps = a.steps
    ncand, nsurv = s_search["n_candidates"], s_search["n_survivors"]
    ncand_tc = s_search["n_candidates_tc"]
    passes = s_build["sort_passes"]  # over all L tables
    rank_path = bool(s_build["rank_path"])
    seg_lists = int(s_search.get("segsort_lists", 0) or 0)   # the last search's hit list was ordered by the segmented sort
    # algorithmic bytes per launch family (DESIGN.md "Kernels"); rank path: u16 bucket ranks + 32-byte
    # fragment records instead of 64-bit packed keys
    rec = ((length + 1) // 2 * 2 + 2 * a.L + 15) // 16 * 16 if rank_path else (length + 15) // 16 * 16
    if rank_path:
        per_table_passes = passes // max(1, a.L)
        hash_bytes = N * (length + 2 * a.L + rec)
        up_bytes = N * 2 * passes
        # first pass: read rank, write (rank, id); every later pass: read and write (rank, id)
        down_bytes = N * a.L * ((2 + 6) + (per_table_passes - 1) * 12)
        group_bytes = N * a.L * 2
        permute_bytes = N * a.L * (4 + rec + length)
    else:
        hash_bytes = N * (length + 8 * KW * a.L)
        up_bytes = N * 8 * passes
        down_bytes = N * ((8 * KW + 4) * 2 * passes - 4 * a.L)
        group_bytes = N * a.L * (8 * KW + 4)
        permute_bytes = N * a.L * (4 + rec + length)
    kern = {
        "hash_fast_kernel": {"ms": acc["ms_hash"] / steps, "bytes": hash_bytes, "launches": 1},
        "radix_downsweep_kernel": {"ms": acc["ms_sort_downsweep"] / steps, "bytes": down_bytes, "launches": passes},
        "radix_upsweep_kernel": {"ms": acc["ms_sort_upsweep"] / steps, "bytes": up_bytes, "launches": passes},
        "bucket_grouping": {"ms": acc["ms_group"] / steps, "bytes": group_bytes, "launches": a.L},
        "gather_blocked_kernel": {"ms": acc["ms_permute"] / steps, "bytes": permute_bytes, "launches": 1},
        "filter_kernel": {"ms": (acc["ms_filter"] - acc["ms_filter_tc"]) / steps,
                          "bytes": (ncand - ncand_tc) * (length + 4) + nsurv * 16, "launches": 1},
        # tensor-core candidate filter (Euclidean metric: filter_mma_kernel): algorithmic flops =
        # 2 * 8*len per (query, member) pair (the <x_m, q> contraction; DESIGN.md "roofline")
        "filter_mma_kernel": {"ms": acc["ms_filter_tc"] / steps, "bytes": ncand_tc * (length + 4), "launches": 1,
                              "flops": 2.0 * ncand_tc * 8 * length},
        # survivor (16) + id (4) + fragment record + hit (24)
        "exact_kernel": {"ms": acc["ms_exact"] / steps, "bytes": nsurv * (16 + 4 + rec) + nh * 24, "launches": 1},
        # radix path: one-word key build (24 + 8), radix passes over (key, perm) and the 24-byte gather;
        # segmented path (csrc/hitsort.cu): bin histogram (24), partition (24 + 12), per-bin sort (2 * 4 + 8 + 24)
        "hit_sort": {"ms": acc["ms_hitsort"] / steps,
                     "bytes": nh * (24 + 36 + 40) if seg_lists else nh * (32 + 6 * 32 + 4 + 48), "launches": 1,
                     "path": "segmented (partition by query + per-bin bucket sort)" if seg_lists else "radix passes"},
    }
    for k, v in kern.items():
        v["gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0
        v["share_of_step"] = v["ms"] / ms_per_step
    dom = max(kern, key=lambda k: kern[k]["ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except (OSError, ValueError):
            traffic = None
    d = kern[dom]
    if "flops" in d:
        # timed inside a long step: the sustained cuBLAS figure is the denominator
        tpeak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
        tsrc = ("measured (MEASURED_PEAKS.json bf16_tflops_sustained; FP16 and BF16 run at the same tensor rate)"
                if "bf16_tflops_sustained" in peaks else "fallback 1590 TFLOP/s")
        ach = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
        roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s",
                    "frac": ach / tpeak, "traffic": traffic, "peak_source": tsrc,
                    "algorithmic_flops_per_launch": d["flops"] / max(1, d["launches"]),
                    "avg_launch_ms": d["ms"] / max(1, d["launches"]), "share_of_step": d["share_of_step"],
                    "note": ("tcgen05 FP16 contraction <x_m, q> over every (query, bucket member) pair, 2*8*len flops "
                             "per pair; ablations (profiles/r02_filter_experiments.md): MMAs 4 ms, TMEM loads 2 ms, scan 9 ms, "
                             "hand-offs between the 22 warps 17 ms of the 35 ms launch; floors: tensor time 6.8 ms, shared-memory "
                             "operand and A-tile traffic 10.5 ms (2.65 GB per SM at 128 B/clk); the tensor pipe is active 22 % of "
                             "the cycles (DESIGN.md, open items)")}
    else:
        roofline = {"kernel": dom, "bound": "hbm", "achieved": d["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": d["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": d["bytes"] / max(1, d["launches"]),
                    "avg_launch_ms": d["ms"] / max(1, d["launches"]), "share_of_step": d["share_of_step"],
                    "note": ""}

    # ---- size-independent checks of the full-size result (outside the timed region) -------
    # rank 0's hits of the last step: reference order, every sampled distance recomputed on the
    # host from the codes (float64, same formula) and inside the threshold
    checks = None
    try:
        from hsearch_b200.capi import HIT_DTYPE
        nv = min(int(nh), cap)
        hv = hits_bufs[(step_no[0] - 1) % nslot][:nv * 24]
        q_t = hv.view(torch.int32)[0::6].to(torch.int64)
        t_t = hv.view(torch.int32)[1::6].to(torch.int64)
        id_t = hv.view(torch.int64)[1::3]
        key = (q_t << 40) | (t_t << 34) | (id_t - rank * N)
        in_order = bool((key[1:] > key[:-1]).all().item()) if nv > 1 else True
        gs = torch.Generator(device=dev)
        gs.manual_seed(7)
        pick = torch.randint(0, max(nv, 1), (min(4096, nv),), device=dev, generator=gs)
        samp = np.frombuffer(hv.view(-1, 24)[pick].cpu().numpy().tobytes(), dtype=HIT_DTYPE)
        xc = codes[torch.from_numpy((samp["db_id"] - rank * N).astype(np.int64)).to(dev)].long()
        xp = table[xc].reshape(len(samp), dim)
        d2 = ((xp - qpts[torch.from_numpy(samp["query"].astype(np.int64)).to(dev)]) ** 2).sum(dim=1).cpu().numpy()
        checks = {"hits_checked_for_order": nv, "reference_order": in_order, "sampled_hits": int(len(samp)),
                  "max_rel_dist2_error": float(np.max(np.abs(d2 - samp["dist2"]) / np.maximum(d2, 1e-300))) if len(samp) else 0.0,
                  "all_within_R": bool(np.all(samp["dist2"] <= a.R * a.R)) if len(samp) else True}
    except Exception as e:  # the checks must never break the bench line
        checks = {"error": repr(e)}

    # ---- exact comparison of a slice of the full-size result with the reference itself ----------
    # Bucket membership is per fragment, so the hits of the full run restricted to an id subset equal
    # the hits of that subset searched alone: the slice db_id < 10^6, query < 10^3 of the last step is
    # compared with the reference's own Search() (oracle/_ref) on those fragments and queries -- order,
    # ids and FP64 distances -- and with the oracle restatement for the first-table column.  A pair the
    # FP16 tensor filter lost would show up here (the same-engine brute force below cannot see it).
    if checks is not None and "error" not in checks and not a.no_subset_check:
        try:
            from oracle.pyoracle import Oracle, Reference
            t0s = time.perf_counter()
            ns, qs = min(N, a.subset_db), min(Q, a.subset_q)
            nv = min(int(nh), cap)
            hv2 = hits_bufs[(step_no[0] - 1) % nslot][:nv * 24].view(-1, 24)
            q_t = hv2.view(torch.int32)[:, 0]
            id_t = hv2.view(torch.int64)[:, 1] - rank * N
            sel = np.ascontiguousarray(hv2[(q_t < qs) & (id_t < ns)].cpu().numpy()).view(HIT_DTYPE).reshape(-1)
            o = Oracle()
            tab = o.coordinates(True)
            db_s = o.embed(codes[:ns].cpu().numpy(), tab)
            qp_s = np.ascontiguousarray(qpts[:qs].cpu().numpy())
            a_s, b_s = o.lsh_tables(12345, dim, a.K, a.L, a.W)
            want, _, _ = o.search(db_s, qp_s, a_s, b_s, a.W, a.R, pred=0, cap=len(sel) + 4096)
            ok_port = bool(len(want) == len(sel) and np.array_equal(sel["query"], want["query"]) and
                           np.array_equal(sel["table_first"], want["table_first"]) and
                           np.array_equal(sel["db_id"], want["db_id"]) and np.array_equal(sel["dist2"], want["dist2"]))
            ok_ref = None
            if Reference.available():
                ref, _, _, _ = Reference().search(db_s, qp_s, a.K, a.L, a.W, a.R, 12345, cap=len(sel) + 4096)
                ok_ref = bool(len(ref) == len(sel) and np.array_equal(sel["query"], ref["query"]) and
                              np.array_equal(sel["db_id"], ref["db_id"]) and np.array_equal(sel["dist2"], ref["dist2"]))
            checks["subset_exact"] = bool(ok_port and ok_ref is not False)
            checks["subset"] = {"db_ids_below": int(ns), "queries_below": int(qs), "hits_in_slice": int(len(sel)),
                                "vs_reference_search": ok_ref, "vs_oracle_port_incl_first_table": ok_port,
                                "reference": "oracle/_ref Search()" if ok_ref is not None else "oracle port only",
                                "cpu_seconds": round(time.perf_counter() - t0s, 1)}
            del db_s, want
        except Exception as e:
            checks["subset_exact"] = None
            checks["subset"] = {"error": repr(e)}

    # ---- recall at full size (outside the timed region; rank 0's shard) ---------------------
    # brute force of all Q x N pairs on the same GPU path (hs_bruteforce_points_dev: tensor filter
    # over the whole DB + exact FP64 stage); every LSH hit is a brute-force hit, so the recall is
    # the ratio of the two lists, unweighted and with the reference's weight()
    # (motif_both_points.cpp:67-87)
    recall = None
    if not a.no_recall:
        try:
            hr = h
            if world > 1:   # every call on a context that joined the communicator is collective: a private one here
                hr = hb.HSearch(length, a.K, a.L, a.W, a.R, table_variant=hb.HS_TABLE_PRINT6, flags=flags, device=local)
                hr.load_fragments_dev(codes.data_ptr(), N, id_base=rank * N)
            with torch.cuda.stream(stream):
                t0r = time.perf_counter()
                nbf = hr.bruteforce_points_dev(qpts.data_ptr(), Q, 0, 0)
                bcap = nbf + 1024
                bf = torch.empty(bcap * 24, dtype=torch.uint8, device=dev)
                nbf = hr.bruteforce_points_dev(qpts.data_ptr(), Q, bf.data_ptr(), bcap)
                torch.cuda.synchronize()
                bf_ms = 1e3 * (time.perf_counter() - t0r) / 2
            s_bf = hr.stats().as_dict()

            lsh_buf = hits_bufs[(step_no[0] - 1) % nslot]
            nl = min(int(nh), cap)
            # evaulate() of the reference on the two device-resident lists (hs_evaluate_recall_dev)
            t0e = time.perf_counter()
            ev = hr.evaluate_recall_dev(bf.data_ptr(), int(nbf), lsh_buf.data_ptr(), nl, Q)
            ev_ms = 1e3 * (time.perf_counter() - t0e)
            nzb = [int(i) for i in ((ev["tp_bin"] + ev["fn_bin"]) > 0).nonzero()[0]]
            deciles = {str(i): round(float(ev["tp_bin"][i]) / float(ev["tp_bin"][i] + ev["fn_bin"][i]), 4)
                       for i in nzb if i % 10 == 9 or i == nzb[-1]}
            recall = {"unweighted": ev["n_tp"] / nbf if nbf else None,
                      "weighted": ev["recall"] if nbf else None,
                      "lsh_hits": nl, "bruteforce_hits": int(nbf), "bruteforce_pairs": int(Q) * int(N),
                      "found_not_in_truth": ev["n_extra"],
                      "recall_by_distance_bin": deciles, "evaluate_wall_ms": round(ev_ms, 2),
                      "bruteforce_ms": round(s_bf["ms_total"], 2), "bruteforce_wall_ms": round(bf_ms, 2),
                      "note": "rank 0's shard; brute force = hs_bruteforce_points_dev (same predicate d2 <= R^2); "
                              "join = hs_evaluate_recall_dev (motif_both_points.cpp:100-165); bins are int(dis*10)"}
            del bf
        except Exception as e:
            recall = {"error": repr(e)}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        cpu = cpu_leg(a, a.cpu_sample)
        cpu.pop("seconds", None)

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f16 tensor filter (f32 accumulate) + f64 exact (hash guard, distances); u16 bucket ranks / u64 keys",
           "data": "synthetic",
           "config": {"workload": workload_name(a), "n_db_per_gpu": N, "n_query": Q, "len": length, "K": a.K,
                      "L": a.L, "W": a.W, "R": a.R, "table": "print6", "sharding": f"db-block x{world}",
                      "hit_gather": ("library path (hs_comm_init / hs_comm_reserve / hs_comm_result): queries broadcast with NCCL; "
                                     "per-(query, table) hit counts all-gathered with NCCL; every rank writes its sorted hits "
                                     "to their final positions of the merged list in rank 0's memory (CUDA IPC mapping, peer "
                                     "stores over NVLink) on a second stream beside the next batch's hash + index build; all "
                                     "merges complete before the timed region ends")
                      if world > 1 else "single rank",
                      "l2": "inputs (>= 1 GB codes, multi-GB keys) exceed the 126 MB L2; no flush needed",
                      "reference_arm": ("--impl reference runs the reference's own Search() on one 10,000-fragment shard per host "
                                        "core against all queries (it stores 640 B per fragment and cannot hold this database); "
                                        "the ratio of the two arms is therefore an extrapolation, and a conservative one")},
           "clocks": clocks, "wall_ms_per_step": wall_ms / a.steps,
           "e2e": e2e, "gpu_launches": int(acc["kernel_launches"]),
           "roofline": roofline, "cpu_baseline": cpu,
           "checks": checks, "multi_gpu_checks": multi, "recall": recall,
           "stages_ms": {k[3:]: round(acc[k] / steps, 4) for k in sorted(acc) if k.startswith("ms_")},
           # gbs = algorithmic bytes / time; for the two filter kernels that figure is nominal (a bucket tile is
           # streamed once for all its queries: the tensor filter's bound is flops, see `roofline`), so it is omitted
           "kernels": {k: ({"ms": round(v["ms"], 4), "share": round(v["share_of_step"], 4)} if k.startswith("filter") else
                           {"ms": round(v["ms"], 4), "gbs": round(v["gbs"], 1), "share": round(v["share_of_step"], 4)})
                       for k, v in kern.items()},
           "counts": {"candidates": int(ncand), "survivors": int(nsurv), "hits_rank0": int(nh),
                      "hits_total": int(nh_total), "query_frags_per_s": Q / ((acc["ms_qhash"] + acc["ms_probe"] +
                                                                             acc["ms_filter"] + acc["ms_exact"] +
                                                                             acc["ms_hitsort"]) / steps * 1e-3),
                      "sort_passes": int(passes), "key_words": int(KW), "rank_path": rank_path,
                      "hit_sort_path": "segmented" if seg_lists else "radix",
                      "guard_hits": int(s_hash["guard_hits"]), "guard_corrected": int(s_hash["guard_corrected"]),
                      "residual_flips": residual_flips}}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cluster(a):
    """configs[3]: all in-bucket pairs within R united (hs_cluster; hclust2.cpp:64-71 pairs,
    union_find.cpp:16-33).  One step = one hs_cluster over the whole DB.  At N > 1 every rank holds the
    DB and the index, the pair work is split and the labels exchanged with NCCL (cluster.cu): total work
    is fixed, "scaling": "strong".  Rank 0 checks the labels of the last step against a host union-find
    over the within-R pairs of a slice (oracle) only at small sizes; at any size the label checksum
    must be the same on every rank and every step."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    import hsearch_b200 as hb
    from hsearch_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = capi.load()
    N, length = a.n_db, a.len
    h = hb.HSearch(length, a.K, a.L, a.W, a.R, table_variant=hb.HS_TABLE_PRINT6, predicate=hb.HS_PRED_SQRT_LE_R,
                   device=local)
    pa, pb = h.seed_projection(12345)
    # families of ten: a root per ten fragments, each member with up to two substituted residues
    g = torch.Generator(device=dev)
    g.manual_seed(5)                      # the same DB on every rank
    nroot = N // 10 + 1
    roots = torch.randint(0, 20, (nroot, length), dtype=torch.uint8, device=dev, generator=g)
    codes = roots[torch.arange(N, device=dev) % nroot].contiguous()
    ar = torch.arange(N, device=dev)
    for _ in range(2):
        m = torch.rand(N, device=dev, generator=g) < 0.5
        pos = torch.randint(0, length, (N,), device=dev, generator=g)
        val = torch.randint(0, 20, (N,), dtype=torch.uint8, device=dev, generator=g)
        codes[ar, pos] = torch.where(m, val, codes[ar, pos])
    del ar, roots
    torch.cuda.synchronize()
    h.load_fragments_dev(codes.data_ptr(), N)
    h.build_index()
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = np.zeros(128, dtype=np.uint8)
            capi.check(lib.hs_comm_unique_id(raw.ctypes.data_as(C.c_void_p)))
            uid = torch.from_numpy(raw)
        uid = uid.to(dev)
        dist.broadcast(uid, 0)
        raw = uid.cpu().numpy()
        capi.check(lib.hs_comm_init(h.ctx, raw.ctypes.data_as(C.c_void_p), rank, world))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    labels = None
    for _ in range(a.warmup):
        labels = h.cluster()
    sync()
    t0w = time.time()
    t0 = time.perf_counter()
    dev_ms, pairs, surv, edges, launches = 0.0, 0, 0, 0, 0
    for _ in range(a.steps):
        labels = h.cluster()
        st = h.stats().as_dict()
        dev_ms += st["ms_total"]
        pairs, surv, edges = st["n_candidates"], st["n_survivors"], st["n_edges"]
        launches += st["kernel_launches"]
    sync()
    t1w = time.time()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop(t0w, t1w) if sampler else None
    # device time of the step (CUDA events of the library around the whole call, label exchange included)
    tt = torch.tensor([dev_ms / a.steps, float(pairs), float(surv), float(edges)], dtype=torch.float64, device=dev)
    csum = int(np.bitwise_xor.reduce((labels.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
                                     + np.arange(len(labels), dtype=np.uint64)))
    same = True
    if world > 1:
        mx = tt[:1].clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt[1:].clone()
        dist.all_reduce(sm)
        ms_step, pairs_all, surv_all, edges_all = float(mx.item()), int(sm[0].item()), int(sm[1].item()), int(sm[2].item())
        cs = torch.tensor([csum & 0x7fffffffffffffff], dtype=torch.int64, device=dev)
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(lo.item() == hi.item())
    else:
        ms_step, pairs_all, surv_all, edges_all = dev_ms / a.steps, pairs, surv, edges
    # Exact check against the oracle on whole families: the components the oracle finds among a sub-sample
    # alone (its own buckets, its own pairs) must each lie inside one component of the full result --
    # every edge of the sub-sample is an edge of the whole database.  (Equality cannot be asked: fragments
    # outside the sub-sample may connect two of its components.)
    sub_check = None
    if rank == 0 and not a.no_subset_check:
        try:
            from oracle.pyoracle import Oracle
            o = Oracle()
            nfam = min(3_000, nroot)   # 30,000 fragments: ~10 s of CPU (the oracle joins every in-bucket pair)
            sub_ids = np.sort((np.arange(nfam)[:, None] + nroot * np.arange(10)[None, :]).reshape(-1))
            sub_ids = sub_ids[sub_ids < N]
            sub_codes = codes[torch.from_numpy(sub_ids).to(dev)].cpu().numpy()
            t0c = time.perf_counter()
            lab_sub, ne_sub = o.cluster(sub_codes, hb.coordinates(hb.HS_TABLE_PRINT6), pa, pb, a.W, a.R, metric=0)
            full_sub = labels[sub_ids]
            ok = bool(np.array_equal(full_sub, full_sub[lab_sub]))
            sub_check = {"fragments": int(len(sub_ids)), "oracle_edges": int(ne_sub),
                         "oracle_components": int(len(np.unique(lab_sub))),
                         "oracle_components_inside_full_components": ok,
                         "full_components_over_the_sample": int(len(np.unique(full_sub))),
                         "cpu_seconds": round(time.perf_counter() - t0c, 1)}
        except Exception as e:  # the check is reported, never silently dropped
            sub_check = {"error": repr(e)}
    if rank == 0:
        out = {"metric": "in_bucket_pairs_joined_per_s", "value": pairs_all / (ms_step * 1e-3), "unit": "pairs/s",
               "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f16 tensor self-join filter (f32 accumulate) + f64 exact; u32 union-find",
               "data": "synthetic",
               "config": {"workload": f"near-pair clustering of {N} synthetic fragments (families of ten), len {length}, "
                                      f"K={a.K} L={a.L} W={a.W:g} R={a.R:g}",
                          "n_db_total": N, "sharding": "replicated DB, pair work split by bucket / bucket chunk; "
                                                       "labels all-gathered (NCCL) and united" if world > 1 else "one GPU",
                          "l2": "the bucket-ordered code stores (L x N x len bytes) exceed the 126 MB L2"},
               "clocks": clocks, "wall_ms_per_step": wall_ms / a.steps, "fragments_per_s": N / (ms_step * 1e-3),
               "gpu_launches": int(launches),
               "counts": {"pairs": pairs_all, "survivors": surv_all, "edges": edges_all,
                          "clusters": int(len(np.unique(labels)))},
               "checks": {"labels_equal_on_all_ranks": same, "label_checksum": f"{csum:016x}",
                          "subsample_vs_oracle": sub_check}}
        print(json.dumps(out), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.workload == "cluster":
        if a.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference's main never reaches UnionFind "
                              "(SURVEY.md 8c); the cluster workload has no reference arm"}))
            return
        d = {"n_db": 20_000_000, "L": 8, "R": 25.0, "steps": 1, "warmup": 1}
        for k, v in d.items():   # configs[3] defaults unless given on the command line
            flag = "--" + k.replace("_", "-")
            if not any(x == flag or x.startswith(flag + "=") for x in sys.argv[1:]):
                setattr(a, k, v)
        run_cluster(a)
        return
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)


if __name__ == "__main__":
    main()
