"""CPU tests of the multi-GPU host logic (hsearch_b200/dist.py) on
the gloo backend, world_size 2: block sharding of the DB, query broadcast, the
hit gather to rank 0 and the reference-order restore.  No GPU, no compute: the
per-rank hit lists are cut from an oracle search so that the gathered + sorted
list can be compared with the single-process answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hsearch_b200 import dist as hdist
from tests.util import hits_as_tuples, planted_queries, random_codes


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [hdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pyoracle import Oracle
        o = Oracle()
        length, K, L, W, R = 10, 4, 4, 50.0, 30.0
        n_total = 6001  # odd: the two shards differ in size
        codes = random_codes(n_total, length, seed=11)
        tab = o.coordinates(True)
        a, b = o.lsh_tables(12345, 8 * length, K, L, W)

        # queries exist on rank 0 only and are broadcast
        if rank == 0:
            q = torch.from_numpy(o.embed(planted_queries(codes, 40, seed=12), tab).copy())
        else:
            q = torch.zeros((40, 8 * length), dtype=torch.float64)
        hdist.broadcast_queries(q, 0)
        qn = q.numpy()

        lo, hi = hdist.shard_range(n_total, rank, world)
        local, _, _ = o.search(o.embed(codes[lo:hi], tab), qn, a, b, W, R)
        local = local.copy()
        local["db_id"] += lo  # id_base of the shard
        buf = torch.from_numpy(local.view(np.uint8).copy())
        if len(local) == 0:
            buf = torch.zeros(0, dtype=torch.uint8)
        gathered, counts = hdist.gather_hits(buf, len(local), 0)
        assert counts[rank] == len(local)
        if rank == 0:
            allh = np.frombuffer(gathered.numpy().tobytes(), dtype=hdist_hit_dtype())
            assert len(allh) == sum(counts)
            got = hdist.sort_hits_reference_order(allh)
            want, _, _ = o.search(o.embed(codes, tab), qn, a, b, W, R)
            # a pair's first-finding table and distance do not depend on the sharding
            assert len(want) > 0
            assert hits_as_tuples(got) == hits_as_tuples(want)
            open(os.path.join(tmpdir, "ok"), "w").write(str(len(want)))
        else:
            assert gathered is None
    finally:
        dist.destroy_process_group()


def hdist_hit_dtype():
    from hsearch_b200.capi import HIT_DTYPE
    assert HIT_DTYPE.itemsize == hdist.HIT_BYTES
    return HIT_DTYPE


@pytest.mark.timeout(180)
def test_gather_hits_world2_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert int(open(tmp_path / "ok").read()) > 0


def _merge_worker(rank, world, port, tmpdir):
    """The library's merge (every rank writes its hits to their final positions of rank 0's list) with
    the transport replaced by gloo: counts all-gathered, positions from dist.merged_positions."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pyoracle import Oracle
        o = Oracle()
        length, K, L, W, R = 10, 4, 4, 50.0, 30.0
        n_total, nq = 9001, 40
        codes = random_codes(n_total, length, seed=3)
        qc = planted_queries(codes, nq, seed=4)
        tab = o.coordinates(True)
        a, b = o.lsh_tables(12345, 8 * length, K, L, W)
        lo, hi = hdist.shard_range(n_total, rank, world)
        local, _, _ = o.search(o.embed(codes[lo:hi], tab), o.embed(qc, tab), a, b, W, R)
        local = local.copy()
        local["db_id"] += lo
        cnt = torch.from_numpy(hdist.segment_counts(local, nq, L))
        allc = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(allc, cnt)
        pos, total = hdist.merged_positions(np.stack([c.numpy() for c in allc]), rank)
        # destination of every local hit: its segment's position + its index inside the segment
        tb = max(1, L.bit_length())
        seg = (local["query"].astype(np.int64) << tb) | local["table_first"].astype(np.int64)
        first = np.concatenate(([0], np.cumsum(cnt.numpy())[:-1]))
        dest = pos[seg] + (np.arange(len(local)) - first[seg])
        lists = [None] * world
        dist.all_gather_object(lists, (dest, local))
        if rank == 0:
            out = np.zeros(total, dtype=local.dtype)
            seen = np.zeros(total, dtype=bool)
            for d, hits in lists:
                assert not seen[d].any()
                out[d] = hits
                seen[d] = True
            assert seen.all()
            want, _, _ = o.search(o.embed(codes, tab), o.embed(qc, tab), a, b, W, R)
            assert len(want) > 0 and hits_as_tuples(out) == hits_as_tuples(want)
            open(os.path.join(tmpdir, "ok"), "w").write(str(len(want)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_merge_positions_world_n_matches_single_process(tmp_path, world):
    mp.spawn(_merge_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert int(open(tmp_path / "ok").read()) > 0


# ---------------------------------------------------------------- sharded cluster (configs[3])
def _cluster_worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pyoracle import Oracle
        from tests.util import planted_families
        import hsearch_b200 as hb
        o = Oracle()
        length, K, L, W, R = 10, 4, 3, 50.0, 25.0
        n_total = 3001
        codes = planted_families(n_total, length, seed=13)
        tab = o.coordinates(True)
        a, b = o.lsh_tables(777, 8 * length, K, L, W)
        lo, hi = hdist.shard_range(n_total, rank, world)
        local = codes[lo:hi]
        buckets = o.hash_codes(local, tab, a, b, W)            # [n_local, L, K]
        strings = o.key_strings(buckets)

        def key_fn(l):
            return np.stack([hb.pack_key_string(s, 1) for s in strings[:, l]])

        def local_edges_fn(l, rc, rg):
            if len(rc) < 2:
                return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
            lab, _ = o.cluster(rc, tab, a[l:l + 1], b[l:l + 1], W, R)
            m = np.nonzero(lab != np.arange(len(lab)))[0]
            return rg[m], rg[lab[m]]

        def union_fn(n, eu, ev):
            return o.union_find_labels(n, eu, ev)

        got = hdist.cluster_sharded(local, lo, n_total, L, key_fn, local_edges_fn, union_fn)
        want, ne = o.cluster(codes, tab, a, b, W, R)
        assert ne > 0
        # labels are component minima in both: compare directly
        assert np.array_equal(got, want[lo:hi])
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write(str(len(np.unique(want))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_cluster_sharded_world2_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_cluster_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    n0, n1 = int(open(tmp_path / "ok0").read()), int(open(tmp_path / "ok1").read())
    assert n0 == n1 and 1 < n0 < 3001


# ---------------------------------------------------------------- sharded recall (R1)
def _recall_worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pyoracle import Oracle
        o = Oracle()
        length, K, L, W, R = 10, 4, 4, 50.0, 30.0
        n_total = 5003
        codes = random_codes(n_total, length, seed=21)
        tab = o.coordinates(True)
        a, b = o.lsh_tables(4242, 8 * length, K, L, W)
        qn = o.embed(planted_queries(codes, 30, seed=22, frac=0.8, max_sub=4), tab)
        lo, hi = hdist.shard_range(n_total, rank, world)
        db = o.embed(codes[lo:hi], tab)
        found, _, _ = o.search(db, qn, a, b, W, R)
        truth = o.bruteforce(db, qn, R, pred=0)
        # the per-shard evaluation (on a GPU: HSearch.evaluate_recall_dev), then one all-reduce
        got = hdist.recall_sharded(o.evaluate(truth, found, R))
        fa, _, _ = o.search(o.embed(codes, tab), qn, a, b, W, R)
        ta = o.bruteforce(o.embed(codes, tab), qn, R, pred=0)
        want = o.evaluate(ta, fa, R)
        assert want["n_tp"] > 0 and want["n_fn"] > 0
        assert (got["n_tp"], got["n_fn"], got["n_extra"]) == (want["n_tp"], want["n_fn"], want["n_extra"])
        assert np.array_equal(got["tp_bin"], want["tp_bin"]) and np.array_equal(got["fn_bin"], want["fn_bin"])
        assert abs(got["tp"] - want["tp"]) <= 1e-12 * want["tp"] and abs(got["fn"] - want["fn"]) <= 1e-12 * want["fn"]
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write(repr(got["recall"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_recall_sharded_world2_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_recall_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = float(open(tmp_path / "ok0").read()), float(open(tmp_path / "ok1").read())
    assert r0 == r1 and 0.0 < r0 < 1.0


def _label_merge_worker(rank, world, port, tmpdir):
    """hs_cluster on a communicator (cluster.cu) with the transport replaced by gloo: every rank
    labels the components of its round-robin share of the edges, the labels are all-gathered and
    merged (dist.merge_partial_labels); every rank must end with the labels of the whole edge set."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(17)               # the same graph on every rank
        n = 4000
        # chains, stars and random edges: components that only close across the ranks' shares
        eu = np.concatenate([np.arange(0, 999), np.full(300, 1500), rng.integers(2000, n, size=900)])
        ev = np.concatenate([np.arange(1, 1000), np.arange(1501, 1801), rng.integers(2000, n, size=900)])
        mine = np.arange(len(eu)) % world == rank
        part = torch.from_numpy(hdist.component_labels(n, eu[mine], ev[mine]).astype(np.int64))
        allp = [torch.zeros_like(part) for _ in range(world)]
        dist.all_gather(allp, part)
        got = hdist.merge_partial_labels(np.stack([p.numpy() for p in allp]))
        want = hdist.component_labels(n, eu, ev)
        assert np.array_equal(got, want)
        assert want[999] == 0 and want[1800] == 1500 and len(np.unique(want)) < n
        assert not np.array_equal(part.numpy(), want)   # a single share does not give the answer
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("1")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_cluster_label_merge_world_n(tmp_path, world):
    mp.spawn(_label_merge_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
