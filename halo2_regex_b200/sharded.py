"""Multi-GPU plumbing for the witness-generation path: one process per GPU, strings sharded by contiguous ranges.

The strings of a batch are independent (reference `derive_states` keeps no cross-string state, src/lib.rs:804-823), so
every rank runs the kernels on its own range and writes its own slice of every witness column; nothing crosses GPUs
except ONE all-reduce (sum) of the lookup multiplicity counters, which are global per table row.
"""
import numpy as np


def shard_plan(offsets, world_size):
    """Contiguous string ranges [lo, hi) per rank, balanced by byte count.  offsets: (N+1,) non-decreasing."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    total = int(offsets[-1] - offsets[0]) if n else 0
    bounds = [0]
    for r in range(1, world_size):
        target = int(offsets[0]) + total * r // world_size
        j = int(np.searchsorted(offsets, np.uint64(target), side="left"))
        bounds.append(min(max(j, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def allreduce_multiplicities(tensors, group=None):
    """Sum the multiplicity counters over all ranks, in place.  `tensors`: the per-def mult / endpoint_mult tensors of a
    DeviceOutputs (views of one flat buffer -> one collective) or any list of int64 tensors (CPU tensors work with gloo)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return
    base = tensors[0]._base if tensors[0]._base is not None else None
    if base is not None and all(t._base is base for t in tensors) and sum(t.numel() for t in tensors) == base.numel():
        dist.all_reduce(base, op=dist.ReduceOp.SUM, group=group)      # the flat buffer of DeviceOutputs: a single collective
    else:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def match_sharded(cfg, data, offsets, rank, world_size, **kw):
    """Host-side convenience: run this rank's contiguous range of a host batch through the C-ABI host entry point and
    all-reduce the multiplicities.  Returns (lo, hi, HostOutputs for strings [lo, hi))."""
    import torch
    lo, hi = shard_plan(offsets, world_size)[rank]
    offs = np.ascontiguousarray(offsets[lo:hi + 1], dtype=np.uint64)
    out, res = cfg.match_batch_host(data, offs, **kw)
    mult = [torch.from_numpy(m.view(np.int64)) for m in out.mult + out.endpoint_mult if m is not None]
    allreduce_multiplicities(mult)
    return lo, hi, out
