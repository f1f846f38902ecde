"""CPU, world_size 2 over gloo: the N>1 host logic — contiguous sharding balanced by bytes and the single all-reduce of
the multiplicity counters.  The compute stand-in is the oracle (test infrastructure); the GPU path is covered by -m gpu."""
import os
import random
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import oracle_config
from test_oracle_golden import SNIPPETS, _random_strings

from halo2_regex_b200.sharded import allreduce_multiplicities, shard_plan


def test_shard_plan_is_contiguous_and_balanced():
    rng = random.Random(1)
    lens = [rng.randrange(0, 300) for _ in range(1000)]
    offs = np.zeros(1001, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    for w in (1, 2, 3, 8):
        plan = shard_plan(offs, w)
        assert plan[0][0] == 0 and plan[-1][1] == 1000
        assert all(plan[i][1] == plan[i + 1][0] for i in range(w - 1))
        sizes = [int(offs[hi] - offs[lo]) for lo, hi in plan]
        assert max(sizes) - min(sizes) <= 2 * 300
    assert shard_plan(np.zeros(1, dtype=np.uint64), 4) == [(0, 0)] * 4
    assert shard_plan(np.array([0, 5], dtype=np.uint64), 3)[-1][1] == 1


def _worker(rank, world, port, strings, M, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = oracle_config("test1", M)
        data = np.frombuffer(b"".join(strings), dtype=np.uint8)
        offs = np.zeros(len(strings) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(s) for s in strings])
        lo, hi = shard_plan(offs, world)[rank]
        out, _ = cfg.match_batch(data, offs[lo:hi + 1])
        mult = [torch.from_numpy(m.view(np.int64)) for m in out.mult + out.endpoint_mult]
        allreduce_multiplicities(mult)
        ret[rank] = (lo, hi, [m.copy() for m in out.mult], [m.copy() for m in out.endpoint_mult], out.masked_chars[:, :M].copy())
    finally:
        dist.destroy_process_group()


def test_two_ranks_match_single_process():
    rng = random.Random(5)
    strings = _random_strings(rng, 300, 120, SNIPPETS)
    M = 121
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, strings, M, ret), nprocs=2, join=True)
    whole, _ = oracle_config("test1", M).match_strings(strings)
    assert ret[0][1] == ret[1][0] and ret[0][0] == 0 and ret[1][1] == len(strings)
    for r in (0, 1):
        lo, hi, mult, emult, mc = ret[r]
        for d in range(2):
            assert np.array_equal(mult[d], whole.mult[d])           # all-reduced: every rank holds the global counters
            assert np.array_equal(emult[d], whole.endpoint_mult[d])
        assert np.array_equal(mc, whole.masked_chars[lo:hi, :M])    # witness columns are just the rank's slice
