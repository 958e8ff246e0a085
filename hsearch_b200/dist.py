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


def sort_hits_reference_order(hits):
    """numpy structured hits -> the reference's output order (query, first table,
    ascending db id; motif_both_points.cpp:224-245)."""
    order = np.lexsort((hits["db_id"], hits["table_first"], hits["query"]))
    return hits[order]
