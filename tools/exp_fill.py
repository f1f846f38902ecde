#!/usr/bin/env python3
"""Experiment (development aid): who should write the 3.5 GB of zeros of config 1?

Times the fused walk kernel (a) as shipped (TMA zero-fill inside), (b) with the fill skipped (B2R_DEBUG=1: timing only, the sparse
columns are wrong), and (b) next to a fill that runs on a side stream: memset kernels (tensor.zero_()) or device-to-device copies
from a small L2-resident zero buffer (cudaMemcpyAsync: copy engines, no SM resources)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch

import halo2_regex_b200 as H
from halo2_regex_b200 import workloads as W
from conftest import product_config

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--set", default="regex1")
ap.add_argument("--pitch", type=int, default=0)
args = ap.parse_args()
N, L = 1 << args.log2n, 1024
M = L + 1
cfg = product_config(args.set, M)
d_bytes = W.config1_torch(N, L, device="cuda").reshape(-1)
d_offs = torch.arange(N + 1, dtype=torch.int64, device="cuda") * L
out = H.DeviceOutputs(cfg, N, compact_pitch=8, max_records=2, row_pitch=args.pitch or None)
algo = N * L + N * 8 + out.written_bytes()
cols = [t for t in [out.masked_chars, out.masked_substr_ids] + list(out.substr_ids) + list(out.start_enable) + list(out.end_enable) if t is not None]
fill_bytes = sum(t.numel() for t in cols)
main = torch.cuda.current_stream()
side = torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(main)
        fn()
        e1.record(main)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def walk():
    cfg.match_batch_device(d_bytes, d_offs, out, stream=main)


def report(name, ms):
    print(f"{name:58s}: {ms:7.3f} ms  -> {algo / ms / 1e6:7.1f} GB/s algorithmic = {algo / ms / 1e6 / 6555.2 * 100:5.1f} % of HBM peak", flush=True)


os.environ.pop("B2R_DEBUG", None)
report(f"fused kernel as shipped (pitch {out.row_pitch})", timed(walk))
os.environ["B2R_DEBUG"] = "1"
t_nofill = timed(walk)
report("fused kernel, zero-fill skipped (timing only)", t_nofill)


def fill_memset(stream):
    with torch.cuda.stream(stream):
        for t in cols:
            t.zero_()


for zmb in (4, 32):
    zsrc = torch.zeros(zmb << 20, dtype=torch.uint8, device="cuda")

    def fill_copy(stream, pieces=1):
        with torch.cuda.stream(stream):
            for t in cols:
                flat = t.view(-1)
                for o in range(0, flat.numel(), zsrc.numel()):
                    k = min(zsrc.numel(), flat.numel() - o)
                    flat[o:o + k].copy_(zsrc[:k], non_blocking=True)

    ms = timed(lambda: fill_copy(main))
    print(f"stand-alone D2D copies from a {zmb} MiB zero buffer: {ms:.3f} ms = {fill_bytes / ms / 1e6:.0f} GB/s written")

    def both_copy():
        side.wait_stream(main)
        fill_copy(side)
        walk()
        main.wait_stream(side)

    report(f"no-fill kernel + D2D zero copies ({zmb} MiB src) on a side stream", timed(both_copy))

ms = timed(lambda: fill_memset(main))
print(f"stand-alone memset of the sparse columns: {ms:.3f} ms = {fill_bytes / ms / 1e6:.0f} GB/s")


def both_memset():
    side.wait_stream(main)
    fill_memset(side)
    walk()
    main.wait_stream(side)


report("no-fill kernel + memset kernels on a side stream", timed(both_memset))


def memset_then_walk():
    fill_memset(main)
    walk()


report("memset, then the no-fill kernel (serial)", timed(memset_then_walk))
os.environ.pop("B2R_DEBUG", None)
