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
