"""Multi-GPU plumbing of the search path, one process per GPU.

The DB fragments shard naturally: rank r owns the contiguous block
[r*N/G, (r+1)*N/G) with global ids id_base + local.  Every rank uses the same
host-generated projection, queries are broadcast from rank 0, each rank probes
and verifies locally, and the hit lists are gathered to rank 0 (the one
exchange step of the path) where a (query, first table, db id) sort restores
the reference's output order.  torch.distributed is the plumbing (NCCL over
NVLink on GPUs, gloo on CPU for the tests); no collective sits on the data path
before the final gather.
"""
import numpy as np
import torch
import torch.distributed as dist

HIT_BYTES = 24


def shard_range(n_total, rank, world):
    """Contiguous block partition; the first n_total % world ranks get one extra."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_queries(q, src=0):
    """q: tensor (device matches the backend).  In place broadcast from src."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(q, src=src)
    return q


def gather_hits(hits_u8, nhits, dst=0):
    """hits_u8: uint8 tensor holding >= nhits*24 bytes of hs_hit records on this
    rank.  Returns (tensor of all ranks' hits, concatenated in rank order, on dst;
    None elsewhere) and the per-rank counts."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return hits_u8[: nhits * HIT_BYTES], [nhits]
    rank = dist.get_rank()
    dev = hits_u8.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([nhits], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine)
    counts = counts.tolist()
    if rank == dst:
        out = torch.empty(sum(counts) * HIT_BYTES, dtype=torch.uint8, device=dev)
        off = 0
        ops = []
        for r in range(world):
            nb = counts[r] * HIT_BYTES
            if r == dst:
                out[off:off + nb].copy_(hits_u8[:nb])
            elif nb:
                ops.append(dist.P2POp(dist.irecv, out[off:off + nb], r))
            off += nb
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out, counts
    nb = nhits * HIT_BYTES
    if nb:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, hits_u8[:nb].contiguous(), dst)]):
            w.wait()
    return None, counts


class HitGather:
    """Gather of the ranks' hit lists to rank `dst`, started asynchronously: the NCCL transfers
    run beside whatever the caller enqueues next (the next batch's hash and index build) and
    wait() completes them.  `out` (on dst) receives the lists concatenated in rank order."""

    def __init__(self, hits_u8, nhits, dst=0, out=None, allgather_pad=None):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.work, self.out, self.counts = [], None, [nhits]
        self.pad = None
        if self.world == 1:
            self.out = hits_u8[: nhits * HIT_BYTES]
            return
        rank = dist.get_rank()
        dev = hits_u8.device
        counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, torch.tensor([nhits], dtype=torch.int64, device=dev))
        self.counts = counts.tolist()
        if allgather_pad is not None:
            # NCCL all-gather of equal-size (padded) blocks: every rank receives all lists; on an
            # NVSwitch box this runs closer to line rate than 7 point-to-point receives into one GPU
            blk = max(self.counts) * HIT_BYTES
            need = blk * self.world
            # the choice between the two transports must be the same on every rank (a rank in the
            # all-gather and another in send/recv would hang): decide on the smallest buffers of all
            sizes = torch.tensor([allgather_pad.numel(), hits_u8.numel()], dtype=torch.int64, device=dev)
            dist.all_reduce(sizes, op=dist.ReduceOp.MIN)
            min_pad, min_hits = sizes.tolist()
            if min_pad >= need and min_hits >= blk:
                self.pad = (allgather_pad[:need], blk)
                self.work = [dist.all_gather_into_tensor(self.pad[0], hits_u8[:blk], async_op=True)]
                self.dst, self.rank = dst, rank
                return
        ops = []
        if rank == dst:
            need = sum(self.counts) * HIT_BYTES
            self.out = out[:need] if out is not None and out.numel() >= need else \
                torch.empty(need, dtype=torch.uint8, device=dev)
            off = 0
            for r in range(self.world):
                nb = self.counts[r] * HIT_BYTES
                if r == dst:
                    self.out[off:off + nb].copy_(hits_u8[:nb], non_blocking=True)
                elif nb:
                    ops.append(dist.P2POp(dist.irecv, self.out[off:off + nb], r))
                off += nb
        else:
            nb = nhits * HIT_BYTES
            if nb:
                ops.append(dist.P2POp(dist.isend, hits_u8[:nb], dst))
        if ops:
            self.work = dist.batch_isend_irecv(ops)

    def wait(self):
        for w in self.work:
            w.wait()
        self.work = []
        if self.pad is not None and self.out is None and self.rank == self.dst:
            buf, blk = self.pad
            # the lists sit at stride blk; rank dst's consumers read them through views
            self.out = [buf[r * blk: r * blk + self.counts[r] * HIT_BYTES] for r in range(self.world)]
        return self.out, self.counts


def segment_counts(hits, n_query, n_tables):
    """Hits in reference order -> count per (query, first table) segment, segment = query * T + table
    with T the table field's width used by the library (comm.cu: 1 << bits(L + 1))."""
    tb = max(1, int(n_tables).bit_length())
    seg = (hits["query"].astype(np.int64) << tb) | hits["table_first"].astype(np.int64)
    return np.bincount(seg, minlength=int(n_query) << tb).astype(np.int64)


def merged_positions(counts_all, rank):
    """The arithmetic of the library's merge on rank 0 (comm.cu, seg_dest_kernel): counts_all[r][s] =
    hits of rank r in segment s.  Ranks own ascending id blocks, so in the merged list (query, first
    table, ascending db id) every segment holds rank 0's hits, then rank 1's, ...  Returns the position
    of the first hit of each of `rank`'s segments and the total number of hits."""
    counts_all = np.asarray(counts_all, dtype=np.int64)
    tot = counts_all.sum(axis=0)
    start = np.concatenate(([0], np.cumsum(tot)[:-1]))
    before = counts_all[:rank].sum(axis=0)
    return start + before, int(tot.sum())


def sort_hits_reference_order(hits):
    """numpy structured hits -> the reference's output order (query, first table,
    ascending db id; motif_both_points.cpp:224-245)."""
    order = np.lexsort((hits["db_id"], hits["table_first"], hits["query"]))
    return hits[order]


# ---- recall of a sharded search (R1 at BASELINE configs[2], 1 B fragments over 8 GPUs) --------
def recall_sharded(local, device="cpu"):
    """evaulate() (motif_both_points.cpp:100-165) of a block-sharded DB: every (query, fragment)
    pair lives on exactly one shard, so the per-shard tp / fn sums, counts and distance bins of
    hs_evaluate_recall(_dev) (HSearch.evaluate_recall*: dict with tp, fn, n_tp, n_fn, n_extra,
    tp_bin, fn_bin) add up; one all-reduce of 2 doubles and 3 + 2*500 counters, no hit list
    leaves its GPU.  Returns the global dict on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    tpb = np.asarray(local["tp_bin"], dtype=np.int64)
    fnb = np.asarray(local["fn_bin"], dtype=np.int64)
    f = torch.tensor([local["tp"], local["fn"]], dtype=torch.float64, device=device)
    c = torch.as_tensor(np.concatenate([[local["n_tp"], local["n_fn"], local["n_extra"]], tpb, fnb]).astype(np.int64),
                        device=device)
    if world > 1:
        dist.all_reduce(f)
        dist.all_reduce(c)
    f, c = f.cpu().numpy(), c.cpu().numpy()
    nb = len(tpb)
    tp, fn = float(f[0]), float(f[1])
    return {"tp": tp, "fn": fn, "recall": tp / (tp + fn) if tp + fn else float("nan"),
            "n_tp": int(c[0]), "n_fn": int(c[1]), "n_extra": int(c[2]),
            "tp_bin": c[3:3 + nb].astype(np.uint64), "fn_bin": c[3 + nb:3 + 2 * nb].astype(np.uint64)}


# ---- sharded near-pair clustering (BASELINE configs[3]) --------------------
# Buckets span shards, so the cluster path has one real exchange step per table: every
# (key, global id, codes) record goes to the rank that owns its bucket (owner = mix(key) mod
# world), the owner finds the in-bucket near pairs of the complete buckets it holds, the
# resulting edges are all-gathered and every rank runs the same union-find over them.
def allgather_rows(t):
    """Variable-length all_gather along dim 0; returns the rows of all ranks in rank order."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return t
    dev = t.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([t.shape[0]], dtype=torch.int64, device=dev))
    counts = counts.tolist()
    mx = max(counts + [1])
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    pad[:t.shape[0]] = t
    out = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx:r * mx + counts[r]] for r in range(world)], dim=0)


def exchange_rows(tensors, dest):
    """Sends row i of every tensor in `tensors` to rank dest[i]; returns the rows this rank
    receives (concatenated in source-rank order).  Grouped point-to-point sends: the
    all-to-all of the cluster path."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return list(tensors)
    rank = dist.get_rank()
    dev = tensors[0].device
    order = torch.argsort(dest, stable=True)
    send_counts = torch.bincount(dest, minlength=world).to(torch.int64)
    recv_counts = torch.zeros(world * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(recv_counts, send_counts.to(dev))
    recv_counts = recv_counts.view(world, world)[:, rank].tolist()   # what each source sends to me
    send_counts = send_counts.tolist()
    outs = []
    for t in tensors:
        ts = t[order].contiguous()
        out = torch.empty((sum(recv_counts),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        ops, so, ro = [], 0, 0
        for r in range(world):
            ns, nr = send_counts[r], recv_counts[r]
            if r == rank:
                out[ro:ro + nr].copy_(ts[so:so + ns])
            else:
                if ns:
                    ops.append(dist.P2POp(dist.isend, ts[so:so + ns], r))
                if nr:
                    ops.append(dist.P2POp(dist.irecv, out[ro:ro + nr], r))
            so += ns
            ro += nr
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        outs.append(out)
    return outs


def bucket_owner(keys, world):
    """keys: int64 tensor [n, KW] (the packed key words reinterpreted); owner rank of each bucket."""
    h = torch.zeros(keys.shape[0], dtype=torch.int64, device=keys.device)
    for w in range(keys.shape[1]):
        h = (h * 1000003) ^ keys[:, w] ^ (keys[:, w] >> 29)
    return (h & 0x7FFFFFFF) % world


def cluster_sharded(codes_local, id_base, n_total, n_tables, key_fn, local_edges_fn, union_fn, device="cpu"):
    """Near-pair clustering of a block-sharded DB.
    key_fn(l) -> uint64 array [n_local, KW]: packed keys of the local fragments in table l.
    local_edges_fn(l, codes [m, len] u8, gids [m] int64) -> (eu, ev) global-id edges among the
    received fragments (complete buckets of table l).
    union_fn(n_total, eu, ev) -> uint32 labels [n_total] (smallest id of each component).
    Returns the labels of the local shard (global ids of the component minima)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    n_local = codes_local.shape[0]
    codes_t = torch.as_tensor(np.ascontiguousarray(codes_local), device=device)
    gids = torch.arange(id_base, id_base + n_local, dtype=torch.int64, device=device)
    eus, evs = [], []
    for l in range(n_tables):
        keys = torch.as_tensor(np.ascontiguousarray(key_fn(l)).view(np.int64), device=device)
        owner = bucket_owner(keys, world)
        rc, rg = exchange_rows([codes_t, gids], owner)
        eu, ev = local_edges_fn(l, rc.cpu().numpy(), rg.cpu().numpy())
        eus.append(np.asarray(eu, dtype=np.int64))
        evs.append(np.asarray(ev, dtype=np.int64))
    e = np.stack([np.concatenate(eus), np.concatenate(evs)], axis=1) if eus else np.zeros((0, 2), dtype=np.int64)
    e_all = allgather_rows(torch.as_tensor(e, device=device)).cpu().numpy()
    labels = union_fn(n_total, e_all[:, 0].astype(np.uint32), e_all[:, 1].astype(np.uint32))
    return labels[id_base:id_base + n_local]


def gpu_cluster_callbacks(h, a, b, codes_local):
    """The three callbacks of cluster_sharded on top of the C ABI: h is an HSearch with the
    local shard loaded and hashed (all L tables); per-table contexts do the in-bucket pair
    search of the received fragments (hs_cluster), hs_union_find merges the edges."""
    from .index import HSearch

    def key_fn(l):
        return h.keys(l)

    def local_edges_fn(l, codes, gids):
        if len(codes) < 2:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
        p = h.params
        with HSearch(h.len, h.K, 1, p.W, p.R, table_variant=p.table_variant, metric=p.metric, predicate=p.predicate,
                     flags=0, device=getattr(h, "device", 0)) as t:
            t.set_projection(a[l:l + 1], b[l:l + 1])
            t.load_fragments(codes)
            t.build_index()
            lab = t.cluster()
        m = np.nonzero(lab != np.arange(len(lab)))[0]
        return gids[m], gids[lab[m]]

    def union_fn(n_total, eu, ev):
        return h.union_find(n_total, eu, ev)

    return key_fn, local_edges_fn, union_fn


# ---- hs_cluster on a communicator (cluster.cu): host mirror of its arithmetic ---------------------
def component_labels(n, eu, ev):
    """label[i] = smallest id of i's component under the edges (eu[k], ev[k]) -- what hs_cluster and
    hs_union_find return (union_find.cpp:16-33: FindRoot / JoinUnion; the partition does not depend
    on the edge order).  Plain numpy label propagation, for tests and small inputs."""
    label = np.arange(n, dtype=np.int64)
    eu = np.asarray(eu, dtype=np.int64)
    ev = np.asarray(ev, dtype=np.int64)
    while True:
        lo = np.minimum(label[eu], label[ev])
        new = label.copy()
        np.minimum.at(new, eu, lo)
        np.minimum.at(new, ev, lo)
        new = new[new]                      # pointer jumping
        if np.array_equal(new, label):
            return label.astype(np.uint32)
        label = new


def merge_partial_labels(all_labels):
    """The exchange step of hs_cluster on several GPUs: every rank labelled the components of ITS share
    of the pairs; label l of fragment i under any rank's share is the edge (i, l) of the whole graph
    (uf_merge_labels_kernel), so the components of all those edges are the components of the whole
    edge set.  all_labels: [world][n] (the all-gathered labels); returns the complete labels."""
    all_labels = np.asarray(all_labels)
    world, n = all_labels.shape
    i = np.tile(np.arange(n, dtype=np.int64), world)
    return component_labels(n, i, all_labels.reshape(-1).astype(np.int64))
