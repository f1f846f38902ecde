#!/usr/bin/env python3
"""End-to-end (host buffers) timing of config 1 with the host-call trace on (development aid): dense, sparse, sparse + reused buffers."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import halo2_regex_b200 as H
import workloads as W
from conftest import product_config

n, L = 1 << 20, 1024
M = L + 1
cfg = product_config("regex1", M)
for k, v in [a.split("=") for a in sys.argv[1:]]:
    cfg.set_option(k, v)
alloc = H.PinnedAllocator()
h_in = alloc(n * L)
h_in[:] = W.config1_torch(n, L, device="cuda").reshape(-1).cpu().numpy()
h_offs = alloc((n + 1) * 8).view(np.uint64)
h_offs[:] = np.arange(n + 1, dtype=np.uint64) * L
hout = H.HostOutputs(n, M, cfg.state_widths, cfg.table_num_rows, cfg.endpoint_num_rows, max_records=2, compact_pitch=8, allocator=alloc)
cfg.match_batch_host(h_in, h_offs, out=hout)
cfg.set_option("trace_host", 1)
modes = (("dense", {}), ("sparse", dict(sparse=True)), ("sparse+reuse", dict(sparse=True, reuse=True)))
if os.environ.get("PROBE_REUSE_ONLY"):
    cfg.match_batch_host(h_in, h_offs, out=hout, sparse=True)
    cfg.match_batch_host(h_in, h_offs, out=hout, sparse=True)
    modes = modes[2:]
for mode, kw in modes:
    for it in range(3):
        t = time.perf_counter()
        cfg.match_batch_host(h_in, h_offs, out=hout, **kw)
        print(f"{mode} call {it}: {(time.perf_counter() - t) * 1e3:.1f} ms, bytes {cfg.last_host_bytes()}", flush=True)
