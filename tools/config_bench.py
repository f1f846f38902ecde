#!/usr/bin/env python3
"""Device-resident timing of BASELINE configs 2 and 4 (development aid; bench.py is the contract and runs config 1)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import halo2_regex_b200 as H
from halo2_regex_b200 import workloads as W
from conftest import product_config

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--set", default="regex3_k3")
ap.add_argument("--log2n", type=int, default=17)
ap.add_argument("--iters", type=int, default=4)
args = ap.parse_args()
base = 1 << 10                                   # strings generated on the host, tiled on the device
if args.config == 2:
    L, M = 1024, 1025
    cfg = product_config(args.set, M)
    data, _ = W.config2_numpy(base, L)
else:
    L, M = 4096, 4097
    allstr, substr, info = W.large_dfa_texts()
    cfg = H.RegexVerifyConfig.configure(M, [H.RegexDefs(H.AllstrRegexDef.read_from_reader(allstr), [H.SubstrRegexDef.read_from_reader(substr)])])
    data, _ = W.config4_numpy(base, L)
N = 1 << args.log2n
d_bytes = torch.from_numpy(data).cuda().repeat(N // base, 1).reshape(-1).contiguous()
d_offs = torch.arange(N + 1, dtype=torch.int64, device="cuda") * L
cfg.set_timing(True)
out = H.DeviceOutputs(cfg, N, compact_pitch=32, max_records=2)
algo = N * L + out.written_bytes()
for it in range(args.iters):
    cfg.match_batch_device(d_bytes, d_offs, out)
    res = cfg.batch_result()
    w, e, f = cfg.last_stage_ms()
    t = w + e + f
    print(f"config {args.config} {args.set if args.config == 2 else 'large DFA'} N=2^{args.log2n}: walk {w:.3f} + emit {e:.3f} + finalize {f:.3f} = {t:.3f} ms -> input {N * L / t / 1e6:.1f} GB/s, "
          f"algorithmic {algo / t / 1e6:.1f} GB/s ({algo / t / 1e6 / 6555.2 * 100:.1f}% of measured HBM peak), plan {cfg.last_plan()}, code {res.code}")
mult = out.mult[0].cpu().numpy().astype(np.uint64)
assert int(mult.sum()) == N * M
