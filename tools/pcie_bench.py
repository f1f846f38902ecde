#!/usr/bin/env python3
"""Pinned host <-> device copy bandwidth of the box (development aid): the ceiling of the end-to-end number in bench.py."""
import time
import torch

n = 1 << 30
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def d2h():
    with torch.cuda.stream(s1):
        h_out.copy_(d_a, non_blocking=True)


def h2d():
    with torch.cuda.stream(s2):
        d_b.copy_(h_in, non_blocking=True)


def both():
    d2h(); h2d()


def d2h_pieces(k=64):
    step = n // k
    with torch.cuda.stream(s1):
        for i in range(k):
            h_out[i * step:(i + 1) * step].copy_(d_a[i * step:(i + 1) * step], non_blocking=True)


print(f"D2H 1 GiB          : {n / timed(d2h) / 1e9:.1f} GB/s")
print(f"H2D 1 GiB          : {n / timed(h2d) / 1e9:.1f} GB/s")
t = timed(both)
print(f"both directions    : {n / t / 1e9:.1f} GB/s each ({2 * n / t / 1e9:.1f} total)")
print(f"D2H in 64 pieces   : {n / timed(d2h_pieces) / 1e9:.1f} GB/s")
