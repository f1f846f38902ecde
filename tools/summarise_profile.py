#!/usr/bin/env python3
"""Turn the ncu outputs of tools/profile_r1.sh (gpurun_out/) into the small tracked summaries under profiles/."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

# ---- launch list of the timed region ------------------------------------------------------------------------------
rows = [r for r in csv.reader(l for l in open(os.path.join(GP, f"{tag}_launches.csv")) if l.startswith('"'))]
hdr = rows[0]
name_i, val_i = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    k = r[name_i]
    t = float(r[val_i].replace(",", ""))
    n, s = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, s + t)
unit = rows[1][hdr.index("Metric Unit")]
tot = sum(s for _, s in agg.values())
with open(os.path.join(OUT, f"{tag}_launches_summary.tsv"), "w") as f:
    f.write(f"# ncu launch list of the timed region of `python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline`\n")
    f.write(f"# (ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and\n")
    f.write(f"# serialised: compare SHARES).  Columns: kernel, launches, total ({unit}), share\n")
    for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:140]}\t{n}\t{s:.1f}\t{100 * s / tot:.1f}%\n")

# ---- full capture of the dominant kernel --------------------------------------------------------------------------
rep = os.path.join(GP, f"{tag}_walk_kernel.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, u, v = rr[0], rr[1], rr[2]
d = dict(zip(h, v))
units = dict(zip(h, u))
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
stall = sorted(k for k in h if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"))
with open(os.path.join(OUT, f"{tag}_walk_kernel_full.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k regex:walk_kernel --launch-skip 3 --launch-count 1, same bench.py command\n")
    for k in want + stall:
        if k in d:
            f.write(f"{k}\t{d[k]}\t{units.get(k, '')}\n")
def num(k):
    return float(d[k].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(units[k], 1)
traffic = {"log2_strings": 20, "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum")}
traffic["dram_bytes_per_launch"] = traffic["dram_bytes_read"] + traffic["dram_bytes_write"]
traffic["source"] = f"ncu --set full, walk_kernel, profiles/{tag}_walk_kernel_full.txt"
json.dump(traffic, open(os.path.join(OUT, f"{tag}_walk_kernel_traffic.json"), "w"), indent=1)
print(open(os.path.join(OUT, f"{tag}_launches_summary.tsv")).read())
print(open(os.path.join(OUT, f"{tag}_walk_kernel_full.txt")).read())
print(traffic)

# ---- where the warps wait: stall samples per SASS instruction (source page) -----------------------------------------
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sr = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(sr) if r and r[0] == "Address")
sh, sd = sr[hi], sr[hi + 1:]
six = {k: i for i, k in enumerate(sh)}
def fnum(r, k):
    try:
        return float(r[six[k]])
    except (ValueError, IndexError, KeyError):
        return 0.0
reasons = [k for k in sh if k.startswith("stall_") and "Not Issued" not in k]
total = sum(fnum(r, "# Samples") for r in sd) or 1.0
with open(os.path.join(OUT, f"{tag}_walk_kernel_stalls.tsv"), "w") as f:
    f.write("# warp-state samples of the full capture by reason, then the 40 SASS instructions holding the most samples\n")
    f.write("# (index = position in the kernel's SASS; a stall is charged to the instruction that could not issue, i.e. the consumer)\n")
    for k in sorted(reasons, key=lambda k: -sum(fnum(r, k) for r in sd)):
        v = sum(fnum(r, k) for r in sd)
        if v / total >= 0.005:
            f.write(f"{k}\t{v:.0f}\t{100 * v / total:.1f}%\n")
    f.write("#\n# index\tsamples\tshare\texecuted\ttop reasons\tinstruction\n")
    top = sorted(range(len(sd)), key=lambda i: -fnum(sd[i], "# Samples"))[:40]
    for i in sorted(top):
        r = sd[i]
        best = sorted(((fnum(r, k), k[6:]) for k in reasons), reverse=True)[:2]
        f.write(f"{i}\t{fnum(r, '# Samples'):.0f}\t{100 * fnum(r, '# Samples') / total:.1f}%\t{fnum(r, 'Instructions Executed'):.0f}\t"
                + ", ".join(f"{n} {v:.0f}" for v, n in best if v) + f"\t{' '.join(r[six['Source']].split())}\n")
print(open(os.path.join(OUT, f"{tag}_walk_kernel_stalls.tsv")).read())
