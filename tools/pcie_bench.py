#!/usr/bin/env python3
"""Pinned host <-> device copy bandwidth of the box: the ceiling of the end-to-end number in bench.py.

    python tools/pcie_bench.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_bench.py   # 8 ranks at once

Every rank copies 1 GiB each way between pinned host memory and its GPU, all ranks at the same time (barrier before, max over
ranks after); rank 0 prints one JSON line with the per-rank and the aggregate rates."""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    if world > 1:
        x = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        dt = float(x.item())
    return dt


def d2h():
    with torch.cuda.stream(s1):
        h_out.copy_(d_a, non_blocking=True)


def h2d():
    with torch.cuda.stream(s2):
        d_b.copy_(h_in, non_blocking=True)


def both():
    d2h(); h2d()


res = {"ranks": world, "bytes_each_way": n, "host_cores": os.cpu_count()}
for name, fn, k in (("d2h", d2h, 1), ("h2d", h2d, 1), ("both", both, 2)):
    t = timed(fn)
    res[name + "_gbs_per_rank"] = n / t / 1e9
    res[name + "_gbs_aggregate"] = k * world * n / t / 1e9
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
