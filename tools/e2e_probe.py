#!/usr/bin/env python3
"""End-to-end (host buffers) timing of config 1 with the host-call trace on (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import halo2_regex_b200 as H
from halo2_regex_b200 import workloads as W
from halo2_regex_b200.buffers import HostOutputs
from conftest import product_config
from bench import pinned_allocator

n, L = 1 << 20, 1024
M = L + 1
cfg = product_config("regex1", M)
alloc, keep = pinned_allocator(torch)
h_in = alloc(n * L)
h_in[:] = W.config1_torch(n, L, device="cuda").reshape(-1).cpu().numpy()
h_offs = alloc((n + 1) * 8).view(np.uint64)
h_offs[:] = np.arange(n + 1, dtype=np.uint64) * L
hout = HostOutputs(n, M, cfg.state_widths, cfg.table_num_rows, cfg.endpoint_num_rows, max_records=2, compact_pitch=8, allocator=alloc)
cfg.match_batch_host(h_in, h_offs, out=hout)
for it in range(3):
    t = time.perf_counter()
    cfg.match_batch_host(h_in, h_offs, out=hout)
    print(f"call {it}: {(time.perf_counter() - t) * 1e3:.1f} ms", flush=True)
