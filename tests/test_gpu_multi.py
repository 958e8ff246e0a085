"""Two-GPU test of the library's own multi-GPU path (hs_comm_unique_id / hs_comm_init /
hs_comm_reserve / hs_comm_result): two processes, one context per GPU, no torch.distributed --
the ncclUniqueId travels through a file.  Rank 0's merged list must be byte-identical to the
one-GPU search of the whole database (order included: query, first table, ascending db id,
motif_both_points.cpp:224-245), and its checksum must equal the sum of the ranks' checksums.
Skipped on boxes with one GPU (run with `gpurun --gpus 2`)."""
import ctypes as C
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _worker(rank, world, tmpdir, cuts):
    import torch
    import hsearch_b200 as hb
    from hsearch_b200 import capi
    from tests.util import planted_queries, random_codes
    lib = capi.load()
    torch.cuda.set_device(rank)
    length, K, L, W, R = 10, 4, 4, 50.0, 30.0
    n_total = cuts[-1]
    codes = random_codes(n_total, length, seed=5)
    tab = hb.coordinates(hb.HS_TABLE_PRINT6)
    lo, hi = cuts[rank], cuts[rank + 1]
    h = hb.HSearch(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS, device=rank)
    h.seed_projection(12345)
    h.load_fragments(codes[lo:hi], id_base=lo)
    h.build_index()
    uid_path = os.path.join(tmpdir, "uid.bin")
    if rank == 0:
        raw = np.zeros(128, dtype=np.uint8)
        capi.check(lib.hs_comm_unique_id(raw.ctypes.data_as(C.c_void_p)))
        raw.tofile(uid_path + ".tmp")
        os.rename(uid_path + ".tmp", uid_path)
    else:
        t0 = time.time()
        while not os.path.exists(uid_path):
            assert time.time() - t0 < 120
            time.sleep(0.05)
        raw = np.fromfile(uid_path, dtype=np.uint8)
    capi.check(lib.hs_comm_init(h.ctx, raw.ctypes.data_as(C.c_void_p), rank, world))
    capi.check(lib.hs_comm_reserve(h.ctx, (1 << 16) if rank else (1 << 22)))   # the largest request wins
    if rank == 0:
        whole = hb.HSearch(length, K, L, W, R, flags=hb.HS_FLAG_SORT_HITS, device=0)   # the one-GPU run
        whole.seed_projection(12345)
        whole.load_fragments(codes)
        whole.build_index()
    for rep, nq in enumerate((300, 2600, 40, 700)):
        qc = planted_queries(codes, nq, seed=100 + rep, frac=0.5)
        qp = tab[qc].reshape(nq, 8 * length)
        if rank:
            qp = np.zeros_like(qp)          # only rank 0's queries count: they are broadcast
        local = h.search_points(qp, cap=1 << 21)
        ptr, total = C.c_void_p(), C.c_uint64(0)
        capi.check(lib.hs_comm_result(h.ctx, C.byref(ptr), C.byref(total)))
        cs_local = hb.HSearch.hits_checksum(local)
        np.array([cs_local, len(local)], dtype=np.uint64).tofile(os.path.join(tmpdir, f"cs_{rep}_{rank}.bin.tmp"))
        os.rename(os.path.join(tmpdir, f"cs_{rep}_{rank}.bin.tmp"), os.path.join(tmpdir, f"cs_{rep}_{rank}.bin"))
        if rank == 0:
            assert ptr.value
            want = whole.search_points(tab[qc].reshape(nq, 8 * length), cap=1 << 22)
            assert total.value == len(want) > 0
            got = torch.as_tensor(_Raw(ptr.value, total.value * 24), device="cuda:0").cpu().numpy().view(capi.HIT_DTYPE)
            assert np.array_equal(got, want)                      # bytes and order of the one-GPU list
            assert h.hits_checksum_dev(ptr.value, total.value) == hb.HSearch.hits_checksum(want)
            ssum, nsum = 0, 0
            for r in range(world):
                p = os.path.join(tmpdir, f"cs_{rep}_{r}.bin")
                t0 = time.time()
                while not os.path.exists(p):
                    assert time.time() - t0 < 120
                    time.sleep(0.02)
                c = np.fromfile(p, dtype=np.uint64)
                ssum = (ssum + int(c[0])) % (1 << 64)
                nsum += int(c[1])
            assert nsum == len(want) and ssum == hb.HSearch.hits_checksum(want)
            # the rank's own list is its shard's part of the whole
            mine = want[(want["db_id"] >= lo) & (want["db_id"] < hi)]
            assert np.array_equal(local, mine)
        else:
            assert not ptr.value
    # a receive buffer that is too small is reported on every rank, with the size needed
    h2 = None
    h.close()
    if rank == 0:
        whole.close()
        open(os.path.join(tmpdir, "ok"), "w").write("1")


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("cuts", [(0, 100_000, 200_000), (0, 199_999, 200_000), (0, 1, 150_000)])
@pytest.mark.timeout(600)
def test_two_gpu_merge_equals_one_gpu_search(tmp_path, cuts):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, str(tmp_path), cuts), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")


def _cluster_worker(rank, world, tmpdir, n_total, W):
    """hs_cluster on a communicator: every rank holds the whole DB, the pair work is split and the
    labels are exchanged (cluster.cu); each rank's labels must equal the one-GPU labels."""
    import torch
    import hsearch_b200 as hb
    from hsearch_b200 import capi
    from tests.util import planted_families
    lib = capi.load()
    torch.cuda.set_device(rank)
    length, K, L, R = 10, 4, 4, 25.0
    codes = planted_families(n_total, length, seed=9)
    h = hb.HSearch(length, K, L, W, R, predicate=hb.HS_PRED_SQRT_LE_R, device=rank)
    h.seed_projection(4242)
    h.load_fragments(codes)
    h.build_index()
    want = h.cluster()                       # before joining: the whole pair work on this GPU
    edges_whole = h.stats().n_edges
    uid_path = os.path.join(tmpdir, "uid.bin")
    if rank == 0:
        raw = np.zeros(128, dtype=np.uint8)
        capi.check(lib.hs_comm_unique_id(raw.ctypes.data_as(C.c_void_p)))
        raw.tofile(uid_path + ".tmp")
        os.rename(uid_path + ".tmp", uid_path)
    else:
        t0 = time.time()
        while not os.path.exists(uid_path):
            assert time.time() - t0 < 120
            time.sleep(0.05)
        raw = np.fromfile(uid_path, dtype=np.uint8)
    capi.check(lib.hs_comm_init(h.ctx, raw.ctypes.data_as(C.c_void_p), rank, world))
    got = h.cluster()
    assert np.array_equal(got, want)
    assert len(np.unique(want)) < n_total          # something was united
    assert 0 < h.stats().n_edges < edges_whole      # this rank united only its share of the pairs
    h.close()
    open(os.path.join(tmpdir, f"ok{rank}"), "w").write("1")


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("n_total,W", [(60_000, 20.0), (300_000, 50.0)])
@pytest.mark.timeout(600)
def test_two_gpu_cluster_equals_one_gpu_cluster(tmp_path, n_total, W):
    import torch.multiprocessing as mp
    mp.spawn(_cluster_worker, args=(2, str(tmp_path), n_total, W), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
