#!/usr/bin/env python3
"""Quick device-resident timing of the walk kernel (development aid; bench.py is the contract)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import halo2_regex_b200 as H
from halo2_regex_b200 import workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=18)
ap.add_argument("--len", type=int, default=1024)
ap.add_argument("--set", default="regex1")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--want", default="")
args = ap.parse_args()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import product_config

N, L = 1 << args.log2n, args.len
M = L + 1
cfg = product_config(args.set, M)
cfg.set_timing(True)
t0 = time.time()
d_bytes = W.config1_torch(N, L, device="cuda").reshape(-1)
torch.cuda.synchronize()
print(f"generated {N} x {L} in {time.time() - t0:.2f}s")
d_offs = torch.arange(N + 1, dtype=torch.int64, device="cuda") * L
want = set(args.want.split(",")) if args.want else None
out = H.DeviceOutputs(cfg, N, compact_pitch=8, max_records=2, want=want)
algo = N * L + out.written_bytes()
for it in range(args.iters):
    cfg.match_batch_device(d_bytes, d_offs, out)
    res = cfg.batch_result()
    w, e, f = cfg.last_stage_ms()
    t = w + e + f
    print(f"iter {it}: walk {w:.3f} + emit {e:.3f} + finalize {f:.3f} = {t:.3f} ms -> input {N * L / t / 1e6:.1f} GB/s, algorithmic {algo / t / 1e6:.1f} GB/s "
          f"({algo / t / 1e6 / 6555.2 * 100:.1f}% of measured HBM peak), plan {cfg.last_plan()}, code {res.code}")

# pure-write reference: zero every sparse column with torch (cudaMemset path)
cols = [t for t in [out.masked_chars, out.masked_substr_ids] + list(out.substr_ids) + list(out.start_enable) + list(out.end_enable) if t is not None]
nbytes = sum(t.numel() * t.element_size() for t in cols)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in cols:
        t.zero_()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"memset of the sparse columns: {nbytes / 1e9:.2f} GB in {ms:.3f} ms = {nbytes / ms / 1e6:.0f} GB/s")
