#!/usr/bin/env python3
"""One-string latency of b2r_match_substrs (development aid): wall clock per call, and the kernels behind it under ncu."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import halo2_regex_b200 as H
from conftest import product_config

cfg = product_config("regex1", 1025)
s = (b"x" * 500 + b"email was meant for @y." + b"z" * 501)[:1024]
for _ in range(20):
    r = cfg.match_substrs(s)
ts = []
for _ in range(200):
    t = time.perf_counter(); r = cfg.match_substrs(s); ts.append(time.perf_counter() - t)
ts.sort()
print(f"match_substrs 1 KiB: p50 {ts[100] * 1e6:.1f} us, p10 {ts[20] * 1e6:.1f}, p99 {ts[198] * 1e6:.1f}; launches {cfg.last_launch_count()}")
cfg.set_option("trace_host", 1)
for _ in range(3):
    cfg.match_substrs(s)
