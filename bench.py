#!/usr/bin/env python3
"""bench.py -- GCUPS of the Smith-Waterman H+P fill (+maxPos, +backtrack) on B200.

    python bench.py --gpus N --steps K --warmup W             # our arm (CUDA, C ABI)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path

A "step" is one pass of the hot path over one synthetic pair: scoring-matrix fill
(H, P), maxPos with the reference tie-break, and the backtrack that negates the path
in P (omp_smithW.c:199-228).  Workload at N=1 = BASELINE.json configs[1]: one
45000 x 45000 pair of generate()-style random DNA (seed 42), int32 H and P, 16.2 GB
written per step.  For N>1 the workload is BASELINE.json configs[2]: ONE 100000 x 100000
pair (80 GB of H+P) in N column strips, one rank per GPU, the strips' fill kernels linked
only by NVLink P2P boundary stores (swb_fill_strip_async); maxPos is an all-gather of one
(score, i, j) triple per rank and the backtrack hops right to left over the strips.
Total work is fixed as N grows, so scaling is "strong".  The replica figure (one
45000 x 45000 pair per GPU, no exchange) and the pair-wise sharded batch (65536 x 256x256,
BASELINE.json configs[4]) are reported as secondary records.

value  = cols*rows / max-over-ranks(step time) / 1e9 with the sequences already in HBM
         (CUDA events on the launching stream, barrier + synchronize on both sides).  At N=1 the K
         timed steps run as a pipeline over three H/P buffer sets: the fills of consecutive steps
         alternate between two streams (the wavefront of step k+1 starts on the SMs the ramp-down of
         step k leaves idle) and the backtrack of step k (a serial pointer chase that occupies one SM)
         runs on a third; nothing is skipped.  `serial` repeats the measurement one step at a time (the latency of a step).
         At N>1 the same pipeline runs over three sets of strip buffers per GPU;
e2e    = the same metric through the host-buffer C-ABI call (swb_ctx_align): H2D of a and
         b, fill, backtrack and the delivery of int32 H and P (16.2 GB) into the caller's pinned host
         buffers inside the timed region.  The library moves one byte per cell over PCIe (row step of
         H + P, swb_pack.cu) and expands it on the host threads; the host buffers are checked against
         the oracle's digests after the timed steps (`e2e.parity`);
roofline = the fill kernel alone: 8 B/cell x (rows+1)(cols+1) cells / its CUDA-event time,
         against MEASURED_PEAKS.json's HBM figure;
cpu_baseline = the unmodified reference (oracle/_ref, built from /root/reference by
         oracle/Makefile) timed on this host's cores on a bounded sub-problem;
         `--impl reference` times ONE FULL 45000 x 45000 pass of it (same configuration as our arm);
parity  = the results of the timed steps against the oracle's committed digests
         (tests/golden/large_digests.json: H and P per row block and column chunk, maxPos, path).

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
    ap.add_argument("--cpu-sample", type=int, default=12288, help="side of the CPU-baseline sub-problem of OUR arm (~10 s)")
    ap.add_argument("--ref-full", type=int, default=1, help="reference arm: time one full pass of the workload (0 = sample only)")
    ap.add_argument("--strip-cols", type=int, default=100000)
    ap.add_argument("--strip-rows", type=int, default=100000)
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary records (batch, score-only, skewed, replicas)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: consecutive fills on ONE stream (no overlap of steps k and k+1)")
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
def time_reference_cpu(side_cols: int, side_rows: int, all_thread_probe: bool = True, published_cfg: bool = False):
    """Times the UNMODIFIED reference program (oracle/_ref/omp_smithW_ref) on this host.
    Returns dict(value GCUPS, cores, kind, sample, seconds).  The reference's per-cell
    `omp critical` (omp_smithW.c:384-387) makes it slower with threads, so both the
    default (all cores) and the 1-thread configuration are probed and the faster one is
    reported -- the most favourable reading of "all the host threads it can use"."""
    from oracle import swo
    ncpu = os.cpu_count() or 1
    notes = []
    extra = {}
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
        if published_cfg and swo.REF_V1_BIN.exists():
            # context: the configuration the reference's authors published (v1, -DSKIP_BACKTRACK=1: no maxPos, no
            # critical section, no backtrack; makefile:9, experiments-lassen/*), all cores
            try:
                pc, pr = min(side_cols, 16384), min(side_rows, 16384)
                tv, _, used = swo.run_reference_cli(pc, pr, threads=None, seed=SEED, binary=swo.REF_V1_BIN, timeout=300.0)
                extra["published_config_v1_skip_backtrack"] = {"value": pc * pr / tv / 1e9, "unit": UNIT, "cores": used or ncpu,
                                                               "sample": f"{pc}x{pr}, fill only, no maxPos/backtrack"}
            except (subprocess.TimeoutExpired, RuntimeError) as e:
                extra["published_config_v1_skip_backtrack"] = {"error": str(e)[:200]}
    else:
        kind = "port"
        orc = swo.Oracle()
        a, b = orc.generate(SEED, side_cols, side_rows)
        t0 = time.perf_counter()
        H, P, mp = orc.fill(a, b, order="wavefront")
        orc.backtrack(P, mp)
        best = dict(sec=time.perf_counter() - t0, cores=1, cols=side_cols, rows=side_rows)
        notes.append("oracle/_ref missing: timed the C restatement (wavefront order) instead")
    gcups = best["cols"] * best["rows"] / best["sec"] / 1e9
    if swo.REF_BIN.exists():
        extra["one_thread_seconds"] = t1_fill + (t1_bt or 0.0)
    out = {"value": gcups, "unit": UNIT, "cores": best["cores"], "kind": kind, "host_cpus": ncpu,
           "sample": f"{best['cols']}x{best['rows']} of the workload, seed {SEED}; " + "; ".join(notes),
           "seconds": best["sec"], "cells": best["cols"] * best["rows"]}
    out.update(extra)
    return out


def host_mem_available_gb() -> float:
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def workload_name(cols, rows, world):
    if world == 1:
        return f"{cols}x{rows} single pair, full int32 H+P fill + maxPos + backtrack"
    return (f"{cols}x{rows} single pair in {world} column strips (one per GPU, NVLink P2P boundary stores), "
            f"full int32 H+P fill + maxPos + backtrack")


def run_reference_arm(args):
    """The reference's own CPU implementation of the path (unmodified omp_smithW.c built by oracle/Makefile) on
    this host's cores, for our arm's metric and config.  N=1: ONE FULL pass of the 45000 x 45000 workload is the
    basis of `value` (same configuration as our arm; ~60 s on one core, 16.2 GB of host memory); the remaining
    steps time a 4096^2 prefix as a cross-check of the rate.  N>1 (100000 x 100000: 80 GB and ~5 min per pass
    on the CPU) times the full 45000 x 45000 pass as a bounded sample of that workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = max(args.gpus, int(os.environ.get("WORLD_SIZE", "1")))
    cols, rows = (args.cols, args.rows) if world == 1 else (args.strip_cols, args.strip_rows)
    full_c, full_r = min(cols, args.cols), min(rows, args.rows)          # the pass that fits a CPU run
    sample = min(4096, args.cpu_sample)
    if args.warmup > 0:                                  # one warm-up pass is enough for a CPU binary
        time_reference_cpu(min(full_c, 1024), min(full_r, 1024), all_thread_probe=False)
    full = None
    need_gb = 8.0 * (full_c + 1) * (full_r + 1) / 1e9 + 4
    if args.ref_full and host_mem_available_gb() > need_gb:
        try:
            full = time_reference_cpu(full_c, full_r, all_thread_probe=True, published_cfg=True)
        except (subprocess.TimeoutExpired, RuntimeError, OSError):
            full = None
    checks = []
    for k in range(max(0, args.steps - (1 if full else 0))):
        r = time_reference_cpu(min(full_c, sample), min(full_r, sample), all_thread_probe=(full is None and k == 0))
        checks.append(r)
    if full is not None and "one_thread_seconds" in full:
        # The basis is the FULL pass on one thread.  The all-thread configuration is only probed at 1024^2, where every
        # anti-diagonal is shorter than CUTOFF = 1024 and the `omp parallel for if(nEle >= CUTOFF)` never forks
        # (omp_smithW.c:209): its rate there is a cache-resident serial rate, not a multi-thread one, and at full size the
        # per-cell `omp critical` (:384-387) makes threads a slowdown (6-12x).  It is reported, not used.
        basis = dict(full)
        basis.update(value=full_c * full_r / full["one_thread_seconds"] / 1e9, cores=1, seconds=full["one_thread_seconds"],
                     cells=full_c * full_r)
        what = f"one full {full_c}x{full_r} pass, 1 thread"
    elif full is not None:
        basis = full
        what = basis["sample"]
    else:
        basis = max(checks, key=lambda r: r["value"])
        what = f"{sample}x{sample} prefix (no full pass: --ref-full 0 or not enough host memory)"
    value = basis["value"]
    same = (world == 1 and basis["cells"] == cols * rows)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cols * rows / value / 1e9 * 1e3, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(cols, rows, world), "seed": SEED, "scoring": [3, -3, -2]},
            "timed": {"basis": what, "seconds": basis["seconds"], "cells": basis["cells"], "same_size_as_workload": same,
                      "cross_check_gcups": [round(r["value"], 5) for r in checks]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": basis["cores"], "kind": basis["kind"],
                             "sample": basis["sample"],
                             **({"published_config_v1_skip_backtrack": basis["published_config_v1_skip_backtrack"]}
                                if "published_config_v1_skip_backtrack" in basis else {})},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- helpers of our arm
KERNELS_PER_FILL = ["prep_kernel", "selector_kernel", "fill_kernel_form<look-up>", "fill_kernel_form<compare> (returns at once)",
                    "argmax_kernel", "finalize_kernel"]


def hbm_peak():
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        return float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def golden_for(cols, rows):
    f = ROOT / "tests" / "golden" / "large_digests.json"
    if not f.exists():
        return None
    return json.loads(f.read_text())["single"].get(f"{cols}x{rows}")


def check_digests(g, Hv, Pv, col0, m_local):
    """Compares the local part of H and P with the oracle's committed digests -> (blocks checked, mismatches)."""
    from oracle.digest import digest_torch
    cols, chunk, B = g["cols"], g["chunk_cols"], g["block_rows"]
    checked = bad = 0
    for key, d in g["blocks"].items():
        i0 = int(key); i1 = min(i0 + B, g["rows"] + 1)
        for c, c0 in enumerate(range(1, cols + 1, chunk)):
            c1 = min(c0 + chunk, cols + 1)
            if c0 < col0 + 1 or c1 > col0 + m_local + 1:
                continue
            checked += 2
            bad += list(digest_torch(Hv[i0:i1, c0 - col0:c1 - col0], i0, c0)) != d["H"][c]
            bad += list(digest_torch(Pv[i0:i1, c0 - col0:c1 - col0], i0, c0)) != d["P"][c]
    return checked, bad


def secondary_single_gpu(swb, torch, dev, local, peak):
    """Kernel-timed records of the other single-GPU BASELINE configurations (not the headline)."""
    out = {}
    stream = torch.cuda.current_stream()
    d_pos = torch.zeros(1, dtype=torch.int64, device=dev); d_sc = torch.zeros(1, dtype=torch.int32, device=dev)
    timer = swb.KernelTimer(local)

    def best_of(fn, reps=3):
        ts = []
        for r in range(reps + 1):
            fn(); torch.cuda.synchronize()
            if r >= 1:
                ts.append(timer.elapsed_ms())
        return min(ts)

    # score only, 45000 x 45000: INT32-pipe roofline.  ops/cell = ALU-pipe instructions per cell in the SASS of the
    # <64,false,true> inner loop (profiles/r02_sass_counts.txt); the ALU pipe issues one warp instruction per 2 clk per
    # SM sub-partition = 16 lanes/clk (tools/ubench.cu, B300_MICROARCH.md "Pipe rates") -> ceiling below
    a, b = swb.generate(SEED, 45000, 45000)
    a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
    ms = best_of(lambda: swb.score_only_async(a_d, 45000, b_d, 45000, 1, d_pos, d_sc, stream=stream, timer=timer))
    sass = ROOT / "profiles" / "r02_sass_counts.json"
    ops = None
    if sass.exists():
        try:
            ops = json.loads(sass.read_text()).get("score_only_alu_ops_per_cell")
        except ValueError:
            ops = None
    rec = {"workload": "45000x45000 score only (no H/P stores)", "kernel_ms": ms, "gcups": 45000 * 45000 / ms / 1e6,
           "maxScore": int(d_sc.item()), "maxPos": int(d_pos.item())}
    if ops:
        lanes = 148 * 4 * 16                       # SMs x sub-partitions x ALU lanes per clock
        ghz = 1.965
        ceil = lanes * ghz / ops                   # G cell updates / s
        rec["roofline_int"] = {"bound": "int32 ALU pipe", "alu_ops_per_cell": ops, "lanes_per_clk": lanes, "sm_ghz": ghz,
                               "ceiling_gcups": ceil, "achieved_gcups": rec["gcups"], "frac": rec["gcups"] / ceil}
    out["score_only"] = rec
    # skewed pair, both orientations (short anti-diagonals: latency bound)
    for (c, r) in ((1000, 2000000), (2000000, 1000)):
        a, b = swb.generate(SEED, c, r)
        a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
        cells = (r + 1) * (c + 1)
        dH = torch.empty(cells, dtype=torch.int32, device=dev); dP = torch.empty(cells, dtype=torch.int32, device=dev)
        ms = best_of(lambda: swb.fill_async(a_d, c, b_d, r, dH, dP, c + 1, d_pos, d_sc, device=local, stream=stream, timer=timer))
        nstrips = (r + 95) // 96                             # single pairs: strips of 96 rows (32 lanes x 3 rows)
        # strips x (steps lane 31 trails lane 0 + poll granularity) + one strip's sweep.  Half skew (lane l trails
        # lane l-1 by two columns) unless the pair is more than three times wider than tall (swb_api.cu: fill_impl)
        skew = 32 if c > 3 * r else 16
        steps_chain = nstrips * (skew + 8) + (c // 4 + skew)
        out[f"skewed_{c}x{r}"] = {"workload": f"{c} cols x {r} rows full fill", "kernel_ms": ms, "gcups": c * r / ms / 1e6,
                                  "hbm_frac": 8.0 * cells / ms / 1e6 / peak, "maxPos": int(d_pos.item()),
                                  "latency_bound": {"chain_steps": steps_chain, "t_step_ns": ms * 1e6 / steps_chain,
                                                    "what": f"strips*({skew}+8) + cols/4+{skew} dependent steps (96-row strips, {'full' if skew == 32 else 'half'} skew); t_step = kernel time / chain steps "
                                                            "(the in-situ cost of one 3x4-cell step when the chain is the only limiter)"}}
        del dH, dP
    # batch: 65536 x 256x256 in one launch (BASELINE configs[4]); sharded pair-wise when N > 1
    out["batch"] = batch_record(swb, torch, dev, local, peak, 0, 65536)
    # several LARGE pairs through the variable-length batch entry point: single-pair kernel, two internal streams
    try:
        c = r = 45000; npairs = 4
        a, b = swb.generate(SEED, c, r)
        A = torch.frombuffer(bytearray(a * npairs), dtype=torch.uint8).to(dev); B = torch.frombuffer(bytearray(b * npairs), dtype=torch.uint8).to(dev)
        size = ((c + 1) * (r + 1) + 3) // 4 * 4
        dH = torch.empty(npairs * size, dtype=torch.int32, device=dev); dP = torch.empty_like(dH)
        pos = torch.zeros(npairs, dtype=torch.int64, device=dev); sc = torch.zeros(npairs, dtype=torch.int32, device=dev)
        offs = lambda step: [k * step for k in range(npairs)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for _ in range(4):
            e0.record(stream)
            swb.fill_pairs_async(A, offs(c), [c] * npairs, B, offs(r), [r] * npairs, offs(size), dH, dP, pos, sc, device=local, stream=stream)
            e1.record(stream); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        out["large_pairs"] = {"workload": f"{npairs} pairs of {c}x{r} through swb_fill_pairs_async (full fill + maxPos per pair; consecutive "
                                          "pairs overlap on two internal streams)",
                              "ms_total": best, "ms_per_pair": best / npairs, "gcups": npairs * c * r / best / 1e6,
                              "hbm_frac": 8.0 * npairs * (c + 1) * (r + 1) / best / 1e6 / peak,
                              "maxPos": sorted(set(int(x) for x in pos.tolist()))}
        del dH, dP
    except Exception as e:
        out["large_pairs"] = {"error": repr(e)[:200]}
    return out


def batch_record(swb, torch, dev, local, peak, first, count):
    import numpy as np
    m = n = 256
    rng = np.random.default_rng(1)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    A = torch.from_numpy(rng.choice(acgt, (65536, m))[first:first + count].copy()).to(dev)
    B = torch.from_numpy(rng.choice(acgt, (65536, n))[first:first + count].copy()).to(dev)
    pitch = m + 1; stride = ((n + 1) * pitch + 3) // 4 * 4
    dH = torch.empty(count * stride, dtype=torch.int32, device=dev); dP = torch.empty(count * stride, dtype=torch.int32, device=dev)
    d_pos = torch.zeros(count, dtype=torch.int64, device=dev); d_sc = torch.zeros(count, dtype=torch.int32, device=dev)
    timer = swb.KernelTimer(local)
    ts = []
    for r in range(4):
        swb.fill_batch_async(A, m, B, n, count, dH, dP, pitch, stride, d_pos, d_sc, device=local,
                             stream=torch.cuda.current_stream(), timer=timer)
        torch.cuda.synchronize()
        if r >= 1:
            ts.append(timer.elapsed_ms())
    ms = min(ts)
    return {"workload": f"{count} pairs of 256x256 (pairs {first}..{first + count - 1} of 65536), full fill, one launch",
            "kernel_ms": ms, "gcups": m * n * count / ms / 1e6, "hbm_frac": 8.0 * count * (n + 1) * pitch / ms / 1e6 / peak,
            "checksum": int(d_pos.sum().item()) ^ int(d_sc.sum().item())}


# --------------------------------------------------------------------------- our arm, N = 1
def run_single(args, torch, swb, dev, local):
    cols, rows = args.cols, args.rows
    cells_padded = (rows + 1) * (cols + 1)
    a, b = swb.generate(SEED, cols, rows)
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

    # ---- parity of what the timed steps produced, against the oracle's committed digests
    parity = {"golden": None}
    g = golden_for(cols, rows)
    if g is not None:
        Pabs = dP.view(rows + 1, cols + 1)
        checked, bad = check_digests(g, dH.view(rows + 1, cols + 1), Pabs.abs(), 0, cols)
        parity = {"golden": "tests/golden/large_digests.json (CPU oracle)", "digests_checked": checked, "digest_mismatches": bad,
                  "maxPos_ok": maxPos == g["maxPos"], "path_len_ok": plen == g["path_len"]}

    # ---- (2) the same K steps as a pipeline: three H/P buffer sets, two fill streams, the backtrack of step k (a serial pointer chase on
    # ONE SM) runs on a high-priority stream beside the fill of step k+1.  Every step still does all of its work; the
    # timed region ends when the last backtrack has finished.  This is the throughput figure (`value`).
    pipelined = not args.no_pipeline
    total_ms = serial_total_ms
    fill_ms_serial = list(fill_ms)
    nsets = 1
    if pipelined:
        try:
            extra = [(torch.empty(cells_padded, dtype=torch.int32, device=dev), torch.empty(cells_padded, dtype=torch.int32, device=dev),
                      torch.zeros(2, dtype=torch.int64, device=dev)) for _ in range(2)]
        except torch.cuda.OutOfMemoryError:
            pipelined = False
    if pipelined:
        # three H/P buffer sets, two fill streams: the fills of consecutive steps alternate between the streams, so the
        # first strips of step k+1 run on the SMs that the ramp-down of step k's wavefront leaves idle (a single fill
        # keeps the 296 strip slots 61 % busy: DESIGN.md section 4); the backtrack of step k runs on a third,
        # high-priority stream.  Set k % 3 is reused by step k+3 after the backtrack of step k.
        sets = [(dH, dP, d_scal)] + extra
        nsets = len(sets)
        s_fills = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        s_fill = s_fills[0]
        s_bt = torch.cuda.Stream(device=dev, priority=-1)
        ptimers = [swb.KernelTimer(local) for _ in range(args.steps)]

        def run_pipeline(nsteps, use_timers):
            e_fill = [torch.cuda.Event() for _ in range(nsteps)]
            e_bt = [torch.cuda.Event() for _ in range(nsteps)]
            for k in range(nsteps):
                H_, P_, sc_ = sets[k % nsets]
                sf = s_fills[k % 2]
                if k >= nsets:
                    sf.wait_event(e_bt[k - nsets])                 # this buffer set is free again
                swb.fill_async(a_d, cols, b_d, rows, H_, P_, cols + 1, sc_[0:1], None, device=local, stream=sf,
                               warps_per_band=args.wpc, timer=ptimers[k] if use_timers else None)
                e_fill[k].record(sf)
                s_bt.wait_event(e_fill[k])
                swb.backtrack_async(P_, cols + 1, d_maxPos=sc_[0:1], d_pathLen=sc_[1:2], device=local, stream=s_bt)
                e_bt[k].record(s_bt)
            for k in range(max(0, nsteps - nsets), nsteps):        # join: the region ends when the last backtracks have
                s_fill.wait_event(e_bt[k])

        run_pipeline(max(args.warmup, 3), False)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        p0.record(s_fill)
        s_fills[1].wait_event(p0)
        run_pipeline(args.steps, True)
        p1.record(s_fill)
        barrier()
        total_ms = p0.elapsed_time(p1)
        fill_ms = [t.elapsed_ms() for t in ptimers]
        for (_, _, sc_) in sets[:min(nsets, args.steps)]:
            assert (int(sc_[0]), int(sc_[1])) == (maxPos, plen), "pipelined steps disagree with the serial ones"
        # the last results of the overlapped region against the oracle's digests (every set that was written)
        if g is not None:
            for (H_, P_, _) in sets[1:min(nsets, args.steps)]:
                c2, b2 = check_digests(g, H_.view(rows + 1, cols + 1), P_.view(rows + 1, cols + 1).abs(), 0, cols)
                parity["digests_checked"] += c2; parity["digest_mismatches"] += b2
        del extra, sets
    # clocks: sampled over a separate, long enough run of the fill (the timed regions above last tens of milliseconds)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    for _ in range(max(20, args.steps)):
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    value = cols * rows / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (the fill): 8 B per cell, HBM-write bound
    peak, peak_src = hbm_peak()
    # the kernel timed ALONE: its CUDA events in the one-at-a-time timed steps (in the pipelined region the backtrack
    # kernel of the previous step holds one SM and some bandwidth beside it; that average is reported next to it)
    fill_avg = statistics.mean(fill_ms_serial)
    achieved = 8.0 * cells_padded / (fill_avg * 1e-3) / 1e9
    traffic, traffic_source = None, None
    tf = ROOT / "profiles" / "fill_traffic.json"            # written from the ncu --set full capture
    if tf.exists():
        try:
            tj = json.loads(tf.read_text())
            if tj.get("cols") == cols and tj.get("rows") == rows:
                traffic = tj.get("dram_bytes_per_launch")
                traffic_source = f"committed ncu --set full capture ({tj.get('source', 'profiles/fill_traffic.json')}), not measured in this run"
        except (ValueError, OSError):
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_source, "kernel": "swb_tall::fill_kernel_form<64,true,true> (96-row strips, half skew, score look-up)",
                "kernel_ms_avg": fill_avg, "kernel_ms_min": min(fill_ms_serial), "kernel_ms_avg_in_pipelined_region": statistics.mean(fill_ms),
                "timed_in": "the K one-at-a-time timed steps (`serial`), CUDA events around the launch on its stream: ONE launch alone on the GPU",
                "sustained": ({"achieved": 8 * cells_padded / (ms_per_step * 1e-3) / 1e9, "frac": 8 * cells_padded / (ms_per_step * 1e-3) / 1e9 / peak,
                               "what": "algorithmic bytes of the K launches / the timed region of the K overlapped steps (everything else in a "
                                       "step included): two launches are in flight at a time, each launch takes longer than ms_per_step "
                                       "(kernel_ms_avg_in_pipelined_region) while the GPU retires one every ms_per_step"} if pipelined else None),
                "algorithmic_bytes_per_launch": 8 * cells_padded,
                "peak_source": peak_src, "fill_gcups": cols * rows / (fill_avg * 1e-3) / 1e9}

    secondary = None
    if not args.no_secondary:
        del dH, dP
        torch.cuda.empty_cache()
        try:
            secondary = secondary_single_gpu(swb, torch, dev, local, peak)
        except Exception as e:                                   # secondary records must never lose the headline
            secondary = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
        dH = dP = None

    # ---- end to end through the host-buffer C ABI (host pinned buffers)
    e2e = None
    if not args.no_e2e:
        dH = dP = None
        torch.cuda.empty_cache()
        nbytes = cells_padded * 4
        hH = swb.host_alloc(nbytes)
        hP = swb.host_alloc(nbytes)
        e2e_parity = None
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
            # what the caller received in HOST memory, against the oracle's digests (uploaded again for the arithmetic)
            if g is not None:
                import ctypes
                import numpy as np
                ncell = (rows + 1) * (cols + 1)
                hHv = np.ctypeslib.as_array(ctypes.cast(hH, ctypes.POINTER(ctypes.c_int32)), shape=(ncell,))
                hPv = np.ctypeslib.as_array(ctypes.cast(hP, ctypes.POINTER(ctypes.c_int32)), shape=(ncell,))
                cH = torch.from_numpy(hHv).to(dev).view(rows + 1, cols + 1)
                cP = torch.from_numpy(hPv).to(dev).view(rows + 1, cols + 1)
                checked, bad = check_digests(g, cH, cP.abs(), 0, cols)
                from oracle.digest import path_digest
                neg = torch.nonzero(cP.view(-1) < 0).view(-1).cpu().numpy()
                e2e_parity = {"digests_checked": checked, "digest_mismatches": bad,
                              "path_cells_ok": bool(neg.size == g["path_len"] and path_digest(neg) == g["path_digest"])}
                cH = cP = None
                torch.cuda.empty_cache()
        finally:
            swb.host_free(hH); swb.host_free(hP)
        packed = os.environ.get("SWB_PACKED_D2H", "") != "0" and nbytes >= (32 << 20)
        d2h = (swb.packed_pitch(cols + 1) * (rows + 1) + 24) if packed else (2 * nbytes + 16)
        e2e = {"value": cols * rows / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": cols + rows,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "host_bytes_delivered_per_step": 2 * nbytes + 16,
               "transfer": ("packed: one byte per cell over PCIe (row step of H + P), expanded into the caller's int32 H and P "
                            f"by {swb.host_threads()} host threads while later chunks are in flight (swb_pack.cu)"
                            if packed else "plain int32 copies"),
               "parity": e2e_parity,
               "api": "swb_ctx_align (host a,b -> host int32 H, P after backtrack, maxPos, path length)"}

    cpu = None
    if not args.no_cpu_baseline:
        cpu = time_reference_cpu(min(cols, args.cpu_sample), min(rows, args.cpu_sample))
        cpu.pop("seconds", None); cpu.pop("cells", None)
    nk = len(KERNELS_PER_FILL) + 1
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(cols, rows, 1),
                       "seed": SEED, "scoring": [3, -3, -2], "pairs": 1,
                       "l2": "each step writes 16.2 GB of H+P (>> 126 MB L2); no flush needed",
                       "parallelism": "1 GPU",
                       "pipeline": ("three H/P buffer sets, two fill streams + one backtrack stream: consecutive steps overlap -- "
                                    "the first strips of step k+1 run on the SMs the ramp-down of step k's wavefront leaves "
                                    "idle, the backtrack of step k (one SM) runs beside them; every step does all of its work "
                                    "inside the timed region, which ends when the last backtrack has finished; `serial` is "
                                    "one step at a time" if pipelined else "none: fill, maxPos, backtrack back to back")},
            "serial": {"ms_per_step": serial_total_ms / args.steps,
                       "value": cols * rows / (serial_total_ms / args.steps * 1e-3) / 1e9,
                       "what": "the same K steps one at a time on one stream (latency of a step)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": nk * args.steps,
            "kernels_per_step": KERNELS_PER_FILL + ["backtrack_kernel"],
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "step_ms": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)},
            "fill_ms": {"min": min(fill_ms), "median": statistics.median(fill_ms)},
            "backtrack_ms_est": statistics.median(step_ms) - statistics.median(fill_ms_serial),
            "result": {"maxPos": maxPos, "path_len": plen}, "secondary": secondary}
    print(json.dumps(line), flush=True)
    return 0


def bind_to_gpu_numa_node(local: int) -> str:
    """Pins this rank to the CPUs closest to its GPU (NVML's CPU affinity of the device) so that the pinned host
    buffers it allocates next are first-touched on that NUMA node: with 8 ranks copying 10 GB each to host memory
    on arbitrary nodes the copies shared the inter-socket links (round 1: an N=8 e2e step took 4.5x an N=1 step)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + bit for w, word in enumerate(words) for bit in range(64) if (word >> bit) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs of GPU {local}'s NUMA node"
        return "NVML affinity outside this container's CPU set: not bound"
    except Exception as e:                                   # affinity is an optimisation, never a failure
        return f"not bound ({type(e).__name__})"


# --------------------------------------------------------------------------- our arm, N > 1: one pair in N column strips
def run_strips(args, torch, dist, swb, dev, local, rank, world):
    strips = importlib.import_module("smith-waterman_b200.strips")
    cols, rows = args.strip_cols, args.strip_rows
    a, b = swb.generate(SEED, cols, rows)
    pipe = strips.StripPipeline(a, b, local)
    st = pipe.strip
    stream = torch.cuda.current_stream()
    timers = [swb.KernelTimer(local) for _ in range(args.steps)]

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    result = {}

    def step(timer=None):
        pipe.fill_async(stream=stream, timer=timer)
        mp = pipe.maxpos()                                 # all-gather of (score, i, j): also separates consecutive fills
        result["maxPos"] = mp
        result["path_len"] = pipe.backtrack(mp)            # right-to-left hops over the strips, one broadcast each

    # ---- (1) one step at a time: fill on every rank, maxPos all-gather, backtrack hops (the latency of a step)
    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    t_host0 = time.perf_counter()
    for k in range(args.steps):
        step(timers[k])
    e1.record(stream)
    barrier()
    t_host = time.perf_counter() - t_host0
    serial_total_ms = e0.elapsed_time(e1)
    total_ms = serial_total_ms
    fill_ms = [t.elapsed_ms() for t in timers]
    fill_ms_serial = list(fill_ms)
    maxPos, plen = result["maxPos"], result["path_len"]

    # ---- (2) the same K steps as a pipeline over three sets of strip buffers (as at N = 1): the maxPos all-gather and the
    # backtrack hops of step k run on a second stream while the fill kernels of step k+1 are already running; every step
    # still does all of its work and the timed region ends when the last backtrack has finished.
    pipelined = not args.no_pipeline
    extra_pipes = []
    if pipelined:
        try:
            extra_pipes = [strips.StripPipeline(a, b, local), strips.StripPipeline(a, b, local)]
        except Exception:                                   # not enough device memory for three buffer sets
            pipelined = False
    ok_t = torch.tensor([1 if pipelined else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    pipelined = bool(ok_t.item())
    overlap_fills = pipelined and not args.no_overlap
    if pipelined:
        # three sets of strip buffers, two fill streams (as at N = 1): the fills of steps k and k+1 overlap on every GPU --
        # a GPU whose strip of step k is done starts its strip of step k+1 while the GPUs to its right still work on
        # step k (in column-strip mode every GPU idles for the hops before and after its own strip) -- and the maxPos
        # all-gather + backtrack hops of step k-1 run on a third stream beside them.
        pipes = [pipe] + extra_pipes
        s_fills = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)] if overlap_fills else [torch.cuda.Stream(device=dev)] * 2
        s_fill = s_fills[0]
        s_bt = torch.cuda.Stream(device=dev, priority=-1)
        ptimers = [swb.KernelTimer(local) for _ in range(args.steps)]
        presult = {}
        lag = 2 if overlap_fills else 1                      # fills enqueued ahead of the step being finished

        def finish(pp, ev):
            with torch.cuda.stream(s_bt):
                s_bt.wait_event(ev)
                mp = pp.maxpos()
                presult["maxPos"] = mp
                presult["path_len"] = pp.backtrack(mp, stream=s_bt)

        def run_pipeline(nsteps, use_timers):
            evs = [torch.cuda.Event() for _ in range(nsteps)]
            for k in range(nsteps):
                # (set k % 3 was last used by step k-3, whose maxPos gather and backtrack finished in iteration k-1)
                pipes[k % 3].fill_async(stream=s_fills[k % 2], timer=ptimers[k] if use_timers else None)
                evs[k].record(s_fills[k % 2])
                if k >= lag:
                    finish(pipes[(k - lag) % 3], evs[k - lag])
            for k in range(max(0, nsteps - lag), nsteps):
                finish(pipes[k % 3], evs[k])

        run_pipeline(max(args.warmup, 3), False)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        p0.record(s_fill)
        s_fills[1].wait_event(p0)
        run_pipeline(args.steps, True)
        p1.record(s_bt)
        barrier()
        total_ms = p0.elapsed_time(p1)
        fill_ms = [t.elapsed_ms() for t in ptimers]
        assert (presult["maxPos"], presult["path_len"]) == (maxPos, plen), "pipelined steps disagree with the serial ones"
        # leave `pipe` holding the result of a complete step (fill + backtrack) for the parity checks below
        step()
        for pp in extra_pipes:
            pp.close()
        # (drop every reference to the extra buffer sets -- 40 GB each at N = 2 -- before the legs below allocate theirs)
        extra_pipes = []; pipes = None; pp = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()

    # the path digest is checked after the last timed backtrack: gather the negated cells of every strip
    neg_local = []
    Pv = st.dP.view(rows + 1, st.pitch)
    for r0 in range(0, rows + 1, 8192):
        idx = torch.nonzero(Pv[r0:r0 + 8192, 1:] < 0)
        if idx.numel():
            neg_local.append((idx[:, 0] + r0) * (cols + 1) + idx[:, 1] + 1 + st.col0)
    neg_local = torch.cat(neg_local) if neg_local else torch.zeros(0, dtype=torch.int64, device=dev)
    cnt = torch.tensor([neg_local.numel()], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    mx = max(int(c.item()) for c in cnts)
    padded = torch.full((max(mx, 1),), -1, dtype=torch.int64, device=dev)
    padded[:neg_local.numel()] = neg_local
    allneg = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(allneg, padded)

    # fill only (no maxPos gather, no backtrack), barrier before each: the kernel-level picture per rank
    fo = []
    for k in range(3):
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        pipe.fill_async(stream=stream)
        f1.record(stream)
        torch.cuda.synchronize()
        fo.append(f0.elapsed_time(f1))
        pipe.maxpos()
    # parity: the local strip against the oracle's digests (chunks of 12500 columns: whole chunks for N = 2, 4, 8)
    g = golden_for(cols, rows)
    checked = bad = 0
    if g is not None:
        Hv = st.dH.view(rows + 1, st.pitch); Pv = st.dP.view(rows + 1, st.pitch)
        checked, bad = check_digests(g, Hv, Pv, st.col0, st.m)
    t = torch.tensor([total_ms, min(fo), statistics.mean(fill_ms), float(checked), float(bad), serial_total_ms], dtype=torch.float64, device=dev)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    kt = torch.tensor([statistics.mean(fill_ms_serial)], dtype=torch.float64, device=dev)      # the kernels of the one-at-a-time steps
    allk = [torch.zeros_like(kt) for _ in range(world)]
    dist.all_gather(allk, kt)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); time.sleep(0.15)
    for _ in range(max(10, args.steps)):
        pipe.fill_async(stream=stream); pipe.maxpos()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    peak, peak_src = hbm_peak()
    ms_per_step = float(tmax[0].item()) / args.steps
    value = cols * rows / (ms_per_step * 1e-3) / 1e9
    kernel_ms = [float(x.item()) for x in allk]
    cells_padded = (rows + 1) * (cols + world)             # every strip stores its own columns plus local column 0
    achieved = 8.0 * cells_padded / (max(kernel_ms) * 1e-3) / 1e9

    # ---- secondary records: replicas (one 45000^2 pair per GPU, no exchange) and the pair-wise sharded batch
    secondary = None
    pipe_closed = False
    if not args.no_secondary:
        try:
            pipe.close(); pipe_closed = True
            del st, Pv
            torch.cuda.empty_cache()
            secondary = {}
            first, count = swb.shard_pairs(65536, world, rank)
            br = batch_record(swb, torch, dev, local, peak, first, count)
            bt = torch.tensor([br["kernel_ms"]], dtype=torch.float64, device=dev)
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
            secondary["batch_sharded"] = {"workload": f"65536 pairs of 256x256 sharded pair-wise over {world} GPUs (contiguous blocks of {count}), no exchange",
                                          "kernel_ms_max_over_ranks": float(bt.item()),
                                          "gcups": 65536 * 256 * 256 / float(bt.item()) / 1e6,
                                          "hbm_frac_per_gpu": br["hbm_frac"]}
            torch.cuda.empty_cache()
            c2, r2 = args.cols, args.rows
            a2, b2 = swb.generate(SEED + rank, c2, r2)
            a2_d = torch.frombuffer(bytearray(a2), dtype=torch.uint8).to(dev); b2_d = torch.frombuffer(bytearray(b2), dtype=torch.uint8).to(dev)
            dH = torch.empty((r2 + 1) * (c2 + 1), dtype=torch.int32, device=dev); dP = torch.empty_like(dH)
            dsc = torch.zeros(2, dtype=torch.int64, device=dev)

            def rstep():
                swb.fill_async(a2_d, c2, b2_d, r2, dH, dP, c2 + 1, dsc[0:1], None, device=local, stream=stream)
                swb.backtrack_async(dP, c2 + 1, d_maxPos=dsc[0:1], d_pathLen=dsc[1:2], device=local, stream=stream)
            rstep(); barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for _ in range(3):
                rstep()
            r1.record(stream); barrier()
            rt = torch.tensor([r0.elapsed_time(r1) / 3], dtype=torch.float64, device=dev)
            dist.all_reduce(rt, op=dist.ReduceOp.MAX)
            secondary["replicas"] = {"workload": f"one {c2}x{r2} pair per GPU (independent pairs, no exchange), fill + maxPos + backtrack, one at a time",
                                     "ms_per_step": float(rt.item()), "gcups": c2 * r2 * world / float(rt.item()) / 1e6, "scaling": "weak"}
            del dH, dP
        except Exception as e:
            secondary = {"error": repr(e)[:300]}
    if not pipe_closed:
        pipe.close()

    # ---- end to end: host a, b -> every rank's strip of H and P in pinned host memory
    e2e = None
    if not args.no_e2e:
        try:
            torch.cuda.empty_cache()
            numa = bind_to_gpu_numa_node(local)
            pipe2 = strips.StripPipeline(a, b, local)
            s2 = pipe2.strip
            nloc = (rows + 1) * s2.pitch
            hH = torch.empty(nloc, dtype=torch.int32).pin_memory(); hP = torch.empty(nloc, dtype=torch.int32).pin_memory()
            a_host = torch.frombuffer(bytearray(a[s2.col0:s2.col0 + s2.m]), dtype=torch.uint8).pin_memory()
            b_host = torch.frombuffer(bytearray(b), dtype=torch.uint8).pin_memory()

            # packed transfer of the strip (swb_pack.cu): columns 1..m_local as one byte per cell + a row base, expanded
            # into the pinned int32 buffers by this rank's share of the host cores; local column 0 (the copy of the
            # left neighbour's last column, P = hand-off marker) goes as it is
            packed = os.environ.get("SWB_PACKED_D2H", "") != "0"
            # host threads per rank: the cores this rank is bound to, shared with the ranks bound to the same set
            mine = tuple(sorted(os.sched_getaffinity(0))) if hasattr(os, "sched_getaffinity") else tuple(range(os.cpu_count() or 1))
            sets = [None] * world
            dist.all_gather_object(sets, mine)
            threads = max(1, len(mine) // max(1, sum(1 for x in sets if x == mine)))
            if packed:
                nscr = swb.d2h_packed_scratch_bytes(rows + 1, s2.m)
                d_scr = torch.empty(nscr, dtype=torch.uint8, device=dev); h_scr = torch.empty(nscr, dtype=torch.uint8).pin_memory()
            hHv, hPv = hH.view(rows + 1, s2.pitch), hP.view(rows + 1, s2.pitch)
            dHv, dPv = s2.dH.view(rows + 1, s2.pitch), s2.dP.view(rows + 1, s2.pitch)

            def estep():
                s2.a_d.copy_(a_host, non_blocking=True); s2.b_d.copy_(b_host, non_blocking=True)
                pipe2.fill_async(stream=stream)
                if not packed:
                    hH.copy_(s2.dH, non_blocking=True)         # H is final after the fill
                mp = pipe2.maxpos()
                pl = pipe2.backtrack(mp)
                if packed:
                    hHv[:, 0].copy_(dHv[:, 0]); hPv[:, 0].copy_(dPv[:, 0])
                    swb.d2h_packed(s2.dH[1:], s2.dP[1:], s2.pitch, rows + 1, s2.m, hH[1:], hP[1:], s2.pitch, d_scr, h_scr,
                                   threads=threads, device=local, stream=stream)
                else:
                    hP.copy_(s2.dP, non_blocking=True)
                torch.cuda.synchronize()
                return mp, pl
            estep(); barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                mp2, pl2 = estep()
            barrier()
            dt = (time.perf_counter() - t0) / args.e2e_steps
            assert (mp2, pl2) == (maxPos, plen)
            te = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dt = float(te.item())
            # what the caller received in HOST memory against the oracle's digests (uploaded again for the arithmetic)
            e2e_parity = None
            if g is not None:
                cH = hH.to(dev).view(rows + 1, s2.pitch); cP = hP.to(dev).view(rows + 1, s2.pitch)
                checked, bad = check_digests(g, cH, cP.abs(), s2.col0, s2.m)
                ok = bool((cH == dHv).all().item() and (cP == dPv).all().item())     # incl. local column 0 and the path marks
                tp = torch.tensor([checked, bad, 0 if ok else 1], dtype=torch.int64, device=dev)
                dist.all_reduce(tp)
                e2e_parity = {"digests_checked": int(tp[0].item()), "digest_mismatches": int(tp[1].item()),
                              "host_equals_device_on_every_rank": int(tp[2].item()) == 0}
                del cH, cP
            d2h_all = torch.tensor([(swb.packed_pitch(s2.m) * (rows + 1) + 12 * (rows + 1) + 24) if packed else (8 * (rows + 1) * s2.pitch + 24)],
                                   dtype=torch.int64, device=dev)
            dist.all_reduce(d2h_all)
            e2e = {"value": cols * rows / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": cols + rows * world,
                   "d2h_bytes_per_step": int(d2h_all.item()), "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
                   "host_bytes_delivered_per_step": 8 * (rows + 1) * (cols + world) + 24 * world,
                   "transfer": (f"packed: one byte per cell over PCIe, expanded into each rank's pinned int32 strip by {threads} host threads per rank"
                                if packed else "plain int32 copies"),
                   "parity": e2e_parity,
                   "host_placement": numa,
                   "api": "StripPipeline (swb_fill_strip_async / swb_backtrack_from_async per rank, swb_d2h_packed): host a, b -> each rank's "
                          "strip of int32 H and P (after backtrack) in pinned host memory, maxPos, path length"}
            pipe2.close()
        except Exception as e:
            e2e = {"error": repr(e)[:300]}

    if rank == 0:
        from oracle.digest import path_digest
        neg = torch.cat([x[x >= 0] for x in allneg]).cpu().numpy()
        parity = {"golden": "tests/golden/large_digests.json (CPU oracle)" if g is not None else None,
                  "digests_checked": int(tsum[3].item()), "digest_mismatches": int(tsum[4].item())}
        if g is not None:
            parity.update(maxPos_ok=maxPos == g["maxPos"], path_len_ok=plen == g["path_len"],
                          path_cells_ok=(neg.size == g["path_len"] and path_digest(neg) == g["path_digest"]))
        nk = len(KERNELS_PER_FILL) + 1
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": workload_name(cols, rows, world), "seed": SEED, "scoring": [3, -3, -2], "pairs": 1,
                           "columns_per_gpu": cols // world,
                           "l2": f"each step writes {8e-9 * cells_padded / world:.1f} GB of H+P per GPU (>> 126 MB L2); no flush needed",
                           "parallelism": f"{world} column strips, one process per GPU; boundary column pushed to the right neighbour with "
                                          "st.relaxed.sys over NVLink inside the fill kernel, per-32-row release flags; NCCL carries only the "
                                          "maxPos all-gather (3 int64 per rank) and one 3-word broadcast per backtrack hop",
                           "timing": "CUDA events on each rank's streams around the K steps, max over ranks; barrier + synchronize on both sides",
                           "pipeline": (("three sets of strip buffers per GPU, two fill streams: the fills of steps k and k+1 overlap (a GPU "
                                         "that has finished its strip of step k starts its strip of step k+1 while the GPUs to its right still "
                                         "work on step k), the maxPos all-gather and the backtrack hops of step k-1 run beside them; "
                                         if overlap_fills else
                                         "three sets of strip buffers per GPU: the maxPos all-gather and the backtrack hops of step k run beside "
                                         "the fill kernels of step k+1; ") +
                                        "every step does all of its work inside the timed region, which ends when the last backtrack has finished"
                                        if pipelined else "none: fill, maxPos, backtrack back to back")},
                "serial": {"ms_per_step": float(tmax[5].item()) / args.steps,
                           "value": cols * rows / (float(tmax[5].item()) / args.steps * 1e-3) / 1e9,
                           "what": "the same K steps one at a time (latency of a step: fill on all ranks, maxPos all-gather, backtrack hops)"},
                "fill_only": {"ms_max_over_ranks": float(tmax[1].item()), "gcups": cols * rows / float(tmax[1].item()) / 1e6,
                              "what": "the strips' fill kernels alone (prep + fill + argmax per rank), barrier before each, best of 3"},
                "kernel_ms_per_rank": [round(x, 3) for x in kernel_ms],
                "host_ms_per_step": t_host / args.steps * 1e3,
                "clocks": clocks, "e2e": e2e, "gpu_launches": nk * args.steps * world,
                "kernels_per_step": KERNELS_PER_FILL + ["backtrack_kernel (on the ranks the path crosses)"],
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                             "traffic": None, "traffic_source": None, "kernel": "swb_tall::fill_kernel_form<64,true,true> (column-strip mode), all ranks",
                             "kernel_ms_avg": max(kernel_ms), "algorithmic_bytes_per_launch": 8 * cells_padded,
                             "peak_source": peak_src + f" x {world} GPUs",
                             "what": "bytes of all strips / slowest rank's fill kernel time / (N x measured HBM peak)"},
                "cpu_baseline": None, "parity": parity,
                "result": {"maxPos": maxPos, "path_len": plen}, "secondary": secondary}
        print(json.dumps(line), flush=True)
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    swb = importlib.import_module("smith-waterman_b200")      # raises if libswb200.so is missing
    if world == 1:
        return run_single(args, torch, swb, dev, local)
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    try:
        return run_strips(args, torch, dist, swb, dev, local, rank, world)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
