#!/usr/bin/env python3
"""BASELINE config 3: a single 64 MiB string through the regex2_test DFA (b2r_match_long).  Development aid."""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import halo2_regex_b200 as H
from conftest import product_config

ap = argparse.ArgumentParser()
ap.add_argument("--log2len", type=int, default=26)
ap.add_argument("--set", default="regex2")
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()
L = 1 << args.log2len
M = L + 1
g = torch.Generator(device="cuda"); g.manual_seed(0xB2000003)
alphabet = torch.tensor([9, 10, 13] + list(range(32, 127)), dtype=torch.uint8, device="cuda")
d = alphabet[torch.randint(0, len(alphabet), (L + 16,), device="cuda", generator=g)][:L].contiguous()
plant = b" Also for xyz."
at = (0xB2000003 * 2654435761) % (L - 64)
d[at:at + len(plant)] = torch.tensor(list(plant), dtype=torch.uint8, device="cuda")
cfg = product_config(args.set, 64)
cfg.set_timing(True)
out = H.DeviceOutputs(cfg, 1, max_records=8, compact_pitch=64, max_chars_size=M)
algo = L + out.written_bytes()
for it in range(args.iters):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cfg.match_long_device(d, out)
    res = cfg.batch_result(check=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    w, e, f = cfg.last_stage_ms()
    print(f"iter {it}: host wall {dt * 1e3:.3f} ms (walk {w:.3f}, zero+emit {e:.3f}, finalize {f:.3f}; prefix stages = the rest) -> input {L / dt / 1e9:.1f} GB/s, "
          f"algorithmic {algo / dt / 1e9:.1f} GB/s, code {res.code}, launches {cfg.last_launch_count()}")
st = out.status.cpu().numpy().view(H._abi.STATUS_DTYPE).reshape(-1)[0]
mc = out.masked_chars[0, at:at + 16].cpu().numpy()
print("status flags", int(st["flags"]), "n_records", int(st["n_records"]), "masked bytes at the planted offset:", bytes(mc))
mult = out.mult[0].cpu().numpy().astype(np.uint64)
assert int(mult.sum()) == M, (int(mult.sum()), M)
print("sum(mult) == M ok")
