#!/usr/bin/env python
"""Host <-> device copy ceiling of the box with all ranks copying at once (torchrun, one rank per GPU):
pinned 1 GiB buffers, H2D alone, D2H alone, and both directions together; prints one JSON line with the
aggregate GB/s.  The end-to-end leg of bench.py moves 1.0 GB in and 1.04 GB out per step and rank."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return reps * n * world * (int(h2d) + int(d2h)) / dt.item() / 1e9


run(True, True, 2)
res = {"ranks": world, "h2d_GBs": round(run(True, False), 1), "d2h_GBs": round(run(False, True), 1),
       "both_GBs": round(run(True, True), 1), "bytes_per_copy": n}
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
