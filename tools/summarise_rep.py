#!/usr/bin/env python3
"""Summarise one `ncu --set full` report of walk_kernel into profiles/<prefix>_full.txt and <prefix>_stalls.tsv.
usage: tools/summarise_rep.py gpurun_out/x.ncu-rep r1_config4 "command that was profiled" """
import csv
import io
import os
import subprocess
import sys

rep, prefix, what = sys.argv[1], sys.argv[2], sys.argv[3]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, u, v = rr[0], rr[1], rr[2]
d, units = dict(zip(h, v)), dict(zip(h, u))
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
stall = sorted(k for k in h if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"))
with open(os.path.join(OUT, f"{prefix}_full.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k regex:walk_kernel, one launch of: {what}\n")
    for k in want + stall:
        if k in d:
            f.write(f"{k}\t{d[k]}\t{units.get(k, '')}\n")
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
with open(os.path.join(OUT, f"{prefix}_stalls.tsv"), "w") as f:
    f.write(f"# warp-state samples by reason, then the 30 SASS instructions holding the most samples; {what}\n")
    for k in sorted(reasons, key=lambda k: -sum(fnum(r, k) for r in sd)):
        val = sum(fnum(r, k) for r in sd)
        if val / total >= 0.005:
            f.write(f"{k}\t{val:.0f}\t{100 * val / total:.1f}%\n")
    f.write("#\n# index\tsamples\tshare\texecuted\ttop reasons\tinstruction\n")
    for i in sorted(sorted(range(len(sd)), key=lambda i: -fnum(sd[i], "# Samples"))[:30]):
        r = sd[i]
        best = sorted(((fnum(r, k), k[6:]) for k in reasons), reverse=True)[:2]
        f.write(f"{i}\t{fnum(r, '# Samples'):.0f}\t{100 * fnum(r, '# Samples') / total:.1f}%\t{fnum(r, 'Instructions Executed'):.0f}\t"
                + ", ".join(f"{n} {x:.0f}" for x, n in best if x) + f"\t{' '.join(r[six['Source']].split())}\n")
print("wrote", prefix)
