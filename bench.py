#!/usr/bin/env python3
"""bench.py — BASELINE.json's metric on BASELINE.json's config, B200 arm and reference (CPU) arm.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle) on the host cores

metric   : input bytes/sec of DFA witness generation (state column, substr ids, enable bitmaps, masked chars/ids,
           status, records, lookup multiplicities) — bit-exact outputs, see tests/.
workload : BASELINE.json configs[1]: test_regexes regex1_test lookup + substr1, 2^20 synthetic 1 KiB strings per GPU,
           max_chars_size M = 1025.  N > 1: every rank owns 2^20 strings of the same global batch
           (weak scaling, strings are independent); the only exchange is an NCCL all-reduce of the multiplicity histograms.
A step   : one pass of the hot path over the rank's batch.  `value` times it with the inputs resident in HBM; `e2e` times
           the reference-facing C-ABI call b2r_match_batch_host with pinned HOST buffers (H2D and D2H inside the region).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
DEFS = os.path.join(ROOT, "tests", "golden", "defs")

LOG2_STRINGS = 20          # per GPU
STRING_LEN = 1024
M = STRING_LEN + 1
METRIC = "input bytes/sec (DFA witness gen)"
UNIT = "GB/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled every ~5 ms through NVML while the timed region runs."""

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.stop_flag, self.t, self.err = gpu_index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:
                try:
                    self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
                except Exception:
                    pass
            time.sleep(0.005)

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        self.stop_flag = True
        self.t.join(timeout=2)
        nv = self.nv
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[1]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": [k for k, v in names.items() if bits & v], "samples": len(sm)}


def load_defs(mod):
    a = os.path.join(DEFS, "regex1_test_lookup.txt")
    s = os.path.join(DEFS, "substr1_test_lookup.txt")
    return a, s


def reference_arm(args):
    """The reference's own CPU algorithm (oracle/oracle.c restatement; the Rust crate cannot be built in this image),
    all host threads, one bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as O
    from halo2_regex_b200 import workloads as W
    a, s = load_defs(O)
    cores = os.cpu_count() or 1
    cfg = O.OracleConfig([(O.OracleAllstr.read_from_text(a), [O.OracleSubstr.read_from_text(s)])], M)
    n = 1 << 14   # 16 MiB of the same synthetic batch per step
    data, _ = W.config1_numpy(n, STRING_LEN)
    offs = np.arange(n + 1, dtype=np.uint64) * STRING_LEN
    out = cfg.new_outputs(n, max_records=2, compact_pitch=8)
    flat = data.reshape(-1)
    for _ in range(args.warmup):
        cfg.match_batch(flat, offs, out=out, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cfg.match_batch(flat, offs, out=out, nthreads=cores)
    dt = time.perf_counter() - t0
    gbs = n * STRING_LEN * args.steps / dt / 1e9
    sample = f"{n} of the 2^{LOG2_STRINGS} strings x {STRING_LEN} B per step, all witness columns + multiplicities"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "regex1_test + substr1, 1 KiB strings, M=1025 (BASELINE configs[1])", "strings_per_step": n, "string_len": STRING_LEN},
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of src/lib.rs:311-888 without halo2 cell assignment / field inversions: faster than the real reference"},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def pinned_allocator(torch):
    keep = []

    def alloc(nbytes):
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
        keep.append(t)
        return t.numpy()[:nbytes]
    return alloc, keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-strings", type=int, default=LOG2_STRINGS, help="strings per GPU (default: the BASELINE config)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time plain launches instead of one CUDA graph per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    # stdout carries exactly one JSON line: everything else that libraries print there (NCCL's version banner) goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist
    import halo2_regex_b200 as H
    from halo2_regex_b200 import workloads as W
    from halo2_regex_b200.buffers import HostOutputs
    from halo2_regex_b200.sharded import allreduce_multiplicities

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # the version banner goes to stdout, which carries the JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    n = 1 << args.log2_strings
    L = STRING_LEN
    a, s = load_defs(H)
    cfg = H.RegexVerifyConfig.configure(M, [H.RegexDefs(H.AllstrRegexDef.read_from_text(a), [H.SubstrRegexDef.read_from_text(s)])], device=local_rank)

    # ---- device-resident arm -----------------------------------------------------------------------------------------
    d_bytes = W.config1_torch(n, L, first=rank * n, device=dev).reshape(-1)      # this rank's slice of the global batch
    d_offs = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    out = H.DeviceOutputs(cfg, n, max_records=2, compact_pitch=8)
    in_bytes = n * L
    algo_bytes = in_bytes + n * 8 + out.written_bytes()     # input + offsets read, every witness column written (M rows/string)
    stream = torch.cuda.current_stream(dev)

    def enqueue():
        cfg.match_batch_device(d_bytes, d_offs, out, stream=torch.cuda.current_stream(dev))
        if world > 1:
            allreduce_multiplicities(out.mult + out.endpoint_mult)            # the path's only exchange (NCCL over NVLink)

    for _ in range(args.warmup):
        enqueue()
    assert cfg.batch_result(stream=stream).code == 0
    launches_per_step = cfg.last_launch_count()
    # One step = a handful of launches (2 memset nodes, walk_kernel, finalize_kernel): on one GPU they are captured in a CUDA
    # graph so that the step is one launch.  (With the NCCL all-reduce inside, the captured step was slower and the process
    # group hung at teardown, so N > 1 times plain launches.)
    graph = None
    if not args.no_graph and world == 1:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                enqueue()
            graph.replay()
            torch.cuda.synchronize()
        except Exception as e:  # pragma: no cover
            print(f"[bench] CUDA graph capture failed ({e!r}); timing plain launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def step():
        if graph is not None:
            graph.replay()
        else:
            enqueue()

    for _ in range(2):
        step()
    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if rank == 0:
        sampler.start()          # before the barrier: NVML start-up on rank 0 must not delay its first step (the others would wait for it)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()      # `ncu --profile-from-start off` lists exactly the launches of the timed region
    t_begin.record(stream)
    for i in range(args.steps):
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
    t_end.record(stream)
    torch.cuda.profiler.stop()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_begin.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    step_ms = sorted(a_.elapsed_time(b_) for a_, b_ in ev)
    assert cfg.batch_result(stream=stream).code == 0

    # size-independent check at full size (the oracle is the checker only in tests/ at small sizes): the all-reduced
    # multiplicities of the last timed step cover every row of every rank
    mult = out.mult[0].cpu().numpy().astype(np.uint64)
    all_rows = n * M * (world if world > 1 else 1)
    assert int(mult.sum()) == all_rows, (int(mult.sum()), all_rows)

    # the dominant kernel alone (walk_kernel with the emit stage fused in), CUDA events on the launching stream inside the library
    cfg.set_timing(True)
    walk, stages = [], []
    for _ in range(min(args.steps, 5)):
        cfg.match_batch_device(d_bytes, d_offs, out, stream=stream)
        cfg.batch_result(stream=stream)
        stages.append(cfg.last_stage_ms())
        walk.append(stages[-1][0])
    cfg.set_timing(False)
    walk_ms = sum(walk) / len(walk)
    plan = cfg.last_plan()

    # ---- end-to-end arm: C-ABI call with pinned host buffers -----------------------------------------------------------
    e2e = None
    if args.e2e_steps > 0:
        # set-up first (5.9 GB of pinned host memory per rank); every rank must succeed before anyone enters the timed part,
        # which contains collectives
        ok = 1
        try:
            alloc, keep = pinned_allocator(torch)
            h_in = alloc(in_bytes)
            h_in[:] = d_bytes.cpu().numpy()
            h_offs = alloc((n + 1) * 8).view(np.uint64)
            h_offs[:] = np.arange(n + 1, dtype=np.uint64) * L
            hout = HostOutputs(n, M, cfg.state_widths, cfg.table_num_rows, cfg.endpoint_num_rows, max_records=2, compact_pitch=8, allocator=alloc)
            cfg.match_batch_host(h_in, h_offs, out=hout)                        # warm-up (device staging allocation)
        except Exception as exc:  # pragma: no cover
            print(f"[bench] end-to-end arm: set-up failed on rank {rank}: {exc!r}", file=sys.stderr)
            ok = 0
        if world > 1:
            t = torch.tensor([ok], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = int(t.item())
        if ok:
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                cfg.match_batch_host(h_in, h_offs, out=hout)
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            d2h = sum(x.nbytes for x in hout.all_arrays())
            e2e = {"value": world * in_bytes * args.e2e_steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": in_bytes + (n + 1) * 8,
                   "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
                   "api": "b2r_match_batch_host (include/b2r.h) with pinned host buffers"}
            assert int(hout.mult[0].sum()) == n * M

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: the oracle, single thread (the reference's threading model), bounded sample ----------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle as O
        ocfg = O.OracleConfig([(O.OracleAllstr.read_from_text(a), [O.OracleSubstr.read_from_text(s)])], M)
        ns = 1 << 16
        data, _ = W.config1_numpy(ns, L)
        offs = np.arange(ns + 1, dtype=np.uint64) * L
        oout = ocfg.new_outputs(ns, max_records=2, compact_pitch=8)
        ocfg.match_batch(data.reshape(-1)[: 1024 * L], offs[:1025], out=ocfg.new_outputs(1024, max_records=2, compact_pitch=8))
        t0 = time.perf_counter()
        ocfg.match_batch(data.reshape(-1), offs, out=oout, nthreads=1)
        dt = time.perf_counter() - t0
        cpu = {"value": ns * L / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port", "host_cores_available": os.cpu_count(),
               "sample": f"first 2^16 of the 2^{args.log2_strings} strings (64 MiB), all witness columns + multiplicities, {dt:.1f} s",
               "note": "C restatement of src/lib.rs:311-888 (hash-map walk, scans) without halo2 cell assignment / field inversions: faster than the real reference"}

    peak, peak_src = measured_peak()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_walk_kernel_traffic.json")) as f:
            tj = json.load(f)
            if tj.get("log2_strings") == args.log2_strings:
                traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    achieved = algo_bytes / (walk_ms * 1e-3) / 1e9
    value = world * in_bytes * args.steps / (total_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "regex1_test + substr1, 2^%d x 1 KiB strings per GPU, M=1025 (BASELINE configs[1])" % args.log2_strings,
                   "strings_per_gpu": n, "string_len": L, "max_chars_size": M, "defs": 1, "states": 29,
                   "l2_policy": "inputs (1 GiB) + outputs (4.6 GB) per step exceed the 126 MB L2; no flush needed",
                   "parallelism": f"strings sharded over {world} GPU(s); NCCL all-reduce of multiplicities only" if world > 1 else "1 GPU",
                   "cuda_graph": graph is not None},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps * world,
        "kernels_per_step": launches_per_step,
        "step_ms_median": step_ms[len(step_ms) // 2],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "walk_kernel<1, u8, bank-replicated tables, shared bins> (DFA walk + fused emit stage: every witness column)",
                     "kernel_ms": walk_ms, "stage_ms": {"walk+emit": walk_ms, "emit_kernel": sum(x[1] for x in stages) / len(stages), "finalize": sum(x[2] for x in stages) / len(stages)},
                     "table_placement": plan[0], "bin_placement": plan[1], "algorithmic_bytes_per_launch": algo_bytes,
                     "bytes_per_input_byte": algo_bytes / in_bytes, "peak_source": peak_src,
                     "frac_of_nominal_8tbs": achieved / 8000.0},     # also against the 8 TB/s spec figure
        "cpu_baseline": cpu,
    }
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
