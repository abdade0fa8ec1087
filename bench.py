#!/usr/bin/env python3
"""bench.py -- GCUPS of the Smith-Waterman H+P fill (+maxPos, +backtrack) on B200.

    python bench.py --gpus N --steps K --warmup W             # our arm (CUDA, C ABI)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path

A "step" is one pass of the hot path over one synthetic pair: scoring-matrix fill
(H, P), maxPos with the reference tie-break, and the backtrack that negates the path
in P (omp_smithW.c:199-228).  Workload at N=1 = BASELINE.json configs[1]: one
45000 x 45000 pair of generate()-style random DNA (seed 42), int32 H and P, 16.2 GB
written per step.  For N>1 every rank runs its own pair (pairs are independent: the
batch workload of the north star is sharded pair-wise, no data-path collective),
so scaling is "weak".

value  = cols*rows*N / max-over-ranks(step time) / 1e9 with the sequences already in HBM
         (CUDA events on the launching stream, barrier + synchronize on both sides).  The K timed
         steps run as a pipeline over two H/P buffer sets: the backtrack of step k (a serial pointer
         chase that occupies one SM) overlaps the fill of step k+1 on a second stream; nothing is
         skipped.  `serial` repeats the measurement one step at a time (the latency of a step);
e2e    = the same metric through the host-buffer C-ABI call (swb_ctx_align): H2D of a and
         b, fill, backtrack and the D2H of H, P (16.2 GB, pinned) inside the timed region;
roofline = the fill kernel alone: 8 B/cell x (rows+1)(cols+1) cells / its CUDA-event time,
         against MEASURED_PEAKS.json's HBM figure;
cpu_baseline = the unmodified reference (oracle/_ref, built from /root/reference by
         oracle/Makefile) timed on this host's cores on a bounded sub-problem.

Only the cpu_baseline / --impl reference legs touch oracle/.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "GCUPS (full H+P fill + maxPos + backtrack)"
UNIT = "GCUPS"
SEED = 42
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cols", type=int, default=45000)
    ap.add_argument("--rows", type=int, default=45000)
    ap.add_argument("--cpu-sample", type=int, default=4096, help="side of the CPU-baseline sub-problem")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="report the one-pair-at-a-time figure as `value`")
    ap.add_argument("--wpc", type=int, default=0, help="warps per band override (0 = library default)")
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples within 25% of the busiest clock seen (idle samples sit at ~120 MHz)
        hi = max(sm)
        load = [x for x in sm if x >= 0.75 * hi]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU arm
def time_reference_cpu(side_cols: int, side_rows: int, all_thread_probe: bool = True):
    """Times the UNMODIFIED reference program (oracle/_ref/omp_smithW_ref) on this host.
    Returns dict(value GCUPS, cores, kind, sample, seconds).  The reference's per-cell
    `omp critical` (omp_smithW.c:384-387) makes it slower with threads, so both the
    default (all cores) and the 1-thread configuration are probed and the faster one is
    reported -- the most favourable reading of "all the host threads it can use"."""
    from oracle import swo
    ncpu = os.cpu_count() or 1
    notes = []
    if swo.REF_BIN.exists():
        kind = "reference"
        t1_fill, t1_bt, _ = swo.run_reference_cli(side_cols, side_rows, threads=1, seed=SEED)
        best = dict(sec=t1_fill + (t1_bt or 0.0), cores=1, cols=side_cols, rows=side_rows)
        notes.append(f"1 thread {side_cols}x{side_rows}: fill {t1_fill:.3f}s + backtrack {t1_bt:.4f}s")
        if all_thread_probe and ncpu > 1:
            pc, pr = min(side_cols, 1024), min(side_rows, 1024)
            try:
                tn_fill, tn_bt, used = swo.run_reference_cli(pc, pr, threads=None, seed=SEED, timeout=60.0)
                rate_n = pc * pr / (tn_fill + (tn_bt or 0.0))
                notes.append(f"default {used} threads {pc}x{pr}: fill {tn_fill:.3f}s")
                if rate_n > best["cols"] * best["rows"] / best["sec"]:
                    best = dict(sec=tn_fill + (tn_bt or 0.0), cores=used or ncpu, cols=pc, rows=pr)
            except subprocess.TimeoutExpired:
                notes.append(f"default {ncpu} threads {pc}x{pr}: >60 s (per-cell omp critical), abandoned")
    else:
        kind = "port"
        import numpy as np
        orc = swo.Oracle()
        a, b = orc.generate(SEED, side_cols, side_rows)
        t0 = time.perf_counter()
        H, P, mp = orc.fill(a, b, order="wavefront")
        orc.backtrack(P, mp)
        best = dict(sec=time.perf_counter() - t0, cores=1, cols=side_cols, rows=side_rows)
        notes.append("oracle/_ref missing: timed the C restatement (wavefront order) instead")
    gcups = best["cols"] * best["rows"] / best["sec"] / 1e9
    return {"value": gcups, "unit": UNIT, "cores": best["cores"], "kind": kind, "host_cpus": ncpu,
            "sample": f"{best['cols']}x{best['rows']} prefix of the workload, seed {SEED}; " + "; ".join(notes),
            "seconds": best["sec"], "cells": best["cols"] * best["rows"]}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    side_c, side_r = min(args.cols, args.cpu_sample), min(args.rows, args.cpu_sample)
    if args.warmup > 0:                                  # one warm-up pass is enough for a CPU binary
        time_reference_cpu(min(side_c, 1024), min(side_r, 1024), all_thread_probe=False)
    res, secs = None, []
    for k in range(args.steps):
        r = time_reference_cpu(side_c, side_r, all_thread_probe=(k == 0))
        secs.append(r["seconds"] / r["cells"])
        if res is None or r["value"] > res["value"]:
            res = r
    per_cell = statistics.mean(secs)
    value = 1.0 / per_cell / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_cell * side_c * side_r * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"{args.cols}x{args.rows} single pair, full H+P fill + backtrack",
                       "timed_sample": f"{side_c}x{side_r} prefix per step (CPU rate is size-independent to ~10%)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                             "sample": res["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- our arm
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    swb = importlib.import_module("smith-waterman_b200")      # raises if libswb200.so is missing

    cols, rows = args.cols, args.rows
    cells_padded = (rows + 1) * (cols + 1)
    a, b = swb.generate(SEED + rank, cols, rows)              # a different pair per rank
    a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev)
    b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
    dH = torch.empty(cells_padded, dtype=torch.int32, device=dev)
    dP = torch.empty(cells_padded, dtype=torch.int32, device=dev)
    d_scal = torch.zeros(2, dtype=torch.int64, device=dev)     # [0] maxPos, [1] path length
    stream = torch.cuda.current_stream()
    timers = [swb.KernelTimer(local) for _ in range(args.steps)]

    def step(timer=None):
        swb.fill_async(a_d, cols, b_d, rows, dH, dP, cols + 1, d_scal[0:1], None, device=local, stream=stream,
                       warps_per_band=args.wpc, timer=timer)
        swb.backtrack_async(dP, cols + 1, d_maxPos=d_scal[0:1], d_pathLen=d_scal[1:2], device=local, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- (1) one pair at a time on one stream: fill, maxPos, backtrack back to back (the latency of a step)
    for _ in range(args.warmup):
        step()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for k in range(args.steps):
        step(timers[k])
        ev[k + 1].record(stream)
    barrier()
    serial_total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    fill_ms = [t.elapsed_ms() for t in timers]
    maxPos, plen = (int(x) for x in d_scal.tolist())

    # ---- (2) the same K steps as a pipeline: two H/P buffer sets, the backtrack of step k (a serial pointer chase on
    # ONE SM) runs on a high-priority stream beside the fill of step k+1.  Every step still does all of its work; the
    # timed region ends when the last backtrack has finished.  This is the throughput figure (`value`).
    pipelined = not args.no_pipeline
    total_ms = serial_total_ms
    fill_ms_serial = list(fill_ms)
    if pipelined:
        try:
            dH2 = torch.empty(cells_padded, dtype=torch.int32, device=dev)
            dP2 = torch.empty(cells_padded, dtype=torch.int32, device=dev)
        except torch.cuda.OutOfMemoryError:
            pipelined = False
    if pipelined:
        sets = [(dH, dP, d_scal), (dH2, dP2, torch.zeros(2, dtype=torch.int64, device=dev))]
        s_fill = torch.cuda.Stream(device=dev)
        s_bt = torch.cuda.Stream(device=dev, priority=-1)
        ptimers = [swb.KernelTimer(local) for _ in range(args.steps)]

        def run_pipeline(nsteps, use_timers):
            e_fill = [torch.cuda.Event() for _ in range(nsteps)]
            e_bt = [torch.cuda.Event() for _ in range(nsteps)]
            for k in range(nsteps):
                H_, P_, sc_ = sets[k % 2]
                if k >= 2:
                    s_fill.wait_event(e_bt[k - 2])                 # this buffer set is free again
                swb.fill_async(a_d, cols, b_d, rows, H_, P_, cols + 1, sc_[0:1], None, device=local, stream=s_fill,
                               warps_per_band=args.wpc, timer=ptimers[k] if use_timers else None)
                e_fill[k].record(s_fill)
                s_bt.wait_event(e_fill[k])
                swb.backtrack_async(P_, cols + 1, d_maxPos=sc_[0:1], d_pathLen=sc_[1:2], device=local, stream=s_bt)
                e_bt[k].record(s_bt)
            for k in range(max(0, nsteps - 2), nsteps):
                s_fill.wait_event(e_bt[k])

        run_pipeline(max(args.warmup, 2), False)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        p0.record(s_fill)
        run_pipeline(args.steps, True)
        p1.record(s_fill)
        barrier()
        total_ms = p0.elapsed_time(p1)
        fill_ms = [t.elapsed_ms() for t in ptimers]
        for (_, _, sc_) in sets[:min(2, args.steps)]:
            assert (int(sc_[0]), int(sc_[1])) == (maxPos, plen), "pipelined steps disagree with the serial ones"
        del dH2, dP2, sets
    # clocks: sampled over a separate, long enough run of the fill (the timed regions above last tens of milliseconds)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    for _ in range(max(20, args.steps)):
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms, serial_total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max, serial_ms_max = float(t[0].item()), float(t[1].item())
    ms_per_step = total_ms_max / args.steps
    value = cols * rows * world / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (the fill): 8 B per cell, HBM-write bound
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    fill_avg = statistics.mean(fill_ms)
    achieved = 8.0 * cells_padded / (fill_avg * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "fill_traffic.json"            # written from the ncu --set full capture
    if tf.exists():
        try:
            tj = json.loads(tf.read_text())
            if tj.get("cols") == cols and tj.get("rows") == rows:
                traffic = tj.get("dram_bytes_per_launch")
        except (ValueError, OSError):
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "swb::fill_kernel", "kernel_ms_avg": fill_avg,
                "kernel_ms_min": min(fill_ms), "algorithmic_bytes_per_launch": 8 * cells_padded,
                "peak_source": peak_src, "fill_gcups": cols * rows / (fill_avg * 1e-3) / 1e9}

    # ---- end to end through the host-buffer C ABI (rank-local pair, host pinned buffers)
    e2e = None
    if not args.no_e2e:
        del dH, dP
        torch.cuda.empty_cache()
        nbytes = cells_padded * 4
        hH = swb.host_alloc(nbytes)
        hP = swb.host_alloc(nbytes)
        try:
            with swb.AlignContext(cols, rows, device=local) as ctx:
                ctx.align(a, b, hH, hP)                      # warm-up (also faults in the pinned pages)
                barrier()
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    mp2, pl2 = ctx.align(a, b, hH, hP)
                barrier()
                dt = (time.perf_counter() - t0) / args.e2e_steps
            assert (mp2, pl2) == (maxPos, plen), "host-buffer call disagrees with the device-resident call"
        finally:
            swb.host_free(hH); swb.host_free(hP)
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        e2e = {"value": cols * rows * world / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": cols + rows,
               "d2h_bytes_per_step": 2 * nbytes + 16, "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "api": "swb_ctx_align (host a,b -> host H, P after backtrack, maxPos, path length)"}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = time_reference_cpu(min(cols, args.cpu_sample), min(rows, args.cpu_sample))
            cpu.pop("seconds", None); cpu.pop("cells", None)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": f"{cols}x{rows} single pair per GPU, full int32 H+P fill + maxPos + backtrack",
                           "seed": SEED, "scoring": [3, -3, -2], "pairs": world,
                           "l2": "each step writes 16.2 GB of H+P (>> 126 MB L2); no flush needed",
                           "parallelism": "pair per GPU, no collective" if world > 1 else "1 GPU",
                           "pipeline": ("two H/P buffer sets per GPU: the backtrack of step k (one SM, high-priority stream) "
                                        "overlaps the fill of step k+1; every step does all of its work inside the timed "
                                        "region" if pipelined else "none: fill, maxPos, backtrack back to back")},
                "serial": {"ms_per_step": serial_ms_max / args.steps,
                           "value": cols * rows * world / (serial_ms_max / args.steps * 1e-3) / 1e9,
                           "what": "the same K steps one at a time on one stream (latency of a step)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": 5 * args.steps,
                "kernels_per_step": ["prep_kernel", "fill_kernel", "argmax_kernel", "finalize_kernel",
                                     "backtrack_kernel"],
                "roofline": roofline, "cpu_baseline": cpu,
                "step_ms": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)},
                "fill_ms": {"min": min(fill_ms), "median": statistics.median(fill_ms)},
                "backtrack_ms_est": statistics.median(step_ms) - statistics.median(fill_ms_serial),
                "result": {"maxPos": maxPos, "path_len": plen}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
