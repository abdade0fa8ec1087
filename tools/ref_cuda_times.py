#!/usr/bin/env python3
"""The reference's own CUDA variants (unmodified, built by oracle/Makefile into oracle/_ref/) timed beside this
library on the same GPU at the sizes SURVEY 8(d) names.  Comparators only: nothing here is part of the product.
  python tools/ref_cuda_times.py [--sizes 2048,8192,25600] > profiles/rNN_ref_cuda.txt
Their 'Elapsed time' is what each program prints: cuda_global_mem_smithW.cu brackets H2D + per-diagonal kernel
launches + D2H (simple-cuda/cuda_global_mem_smithW.cu:382-437); sw-rotated.cu brackets its smithWaterman() call."""
import argparse, importlib, re, subprocess, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
swb = importlib.import_module("smith-waterman_b200")
ap = argparse.ArgumentParser(); ap.add_argument("--sizes", default="2048,8192,25600"); ap.add_argument("--timeout", type=float, default=240.0)
args = ap.parse_args()
dev = torch.device("cuda:0")
print(f"{'size':>7} | {'ours: fill kernel':>18} | {'ours: host->host call':>22} | {'simple-cuda (ms)':>17} | {'rotated-cuda (ms)':>17}")
for n in (int(x) for x in args.sizes.split(",")):
    a, b = swb.generate(42, n, n)
    a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
    dH = torch.empty((n + 1) * (n + 1), dtype=torch.int32, device=dev); dP = torch.empty_like(dH)
    timer = swb.KernelTimer(0); ks = []
    for _ in range(5):
        swb.fill_async(a_d, n, b_d, n, dH, dP, n + 1, None, None, stream=torch.cuda.current_stream(), timer=timer)
        torch.cuda.synchronize(); ks.append(timer.elapsed_ms())
    del dH, dP
    nbytes = (n + 1) * (n + 1) * 4
    hH, hP = swb.host_alloc(nbytes), swb.host_alloc(nbytes)
    with swb.AlignContext(n, n, device=0) as ctx:
        ctx.align(a, b, hH, hP, do_backtrack=False)
        t0 = time.perf_counter(); ctx.align(a, b, hH, hP, do_backtrack=False); e2e = (time.perf_counter() - t0) * 1e3
    swb.host_free(hH); swb.host_free(hP)
    ref = {}
    for name in ("simple_cuda_ref", "rotated_cuda_ref"):
        exe = ROOT / "oracle" / "_ref" / name
        if not exe.exists():
            ref[name] = "not built"; continue
        try:
            out = subprocess.run([str(exe), str(n), str(n)], capture_output=True, text=True, timeout=args.timeout)
            mm = re.search(r"Elapsed time:\s*([0-9.]+)\s*ms", out.stdout)
            ref[name] = mm.group(1) if mm else f"rc={out.returncode}"
        except subprocess.TimeoutExpired:
            ref[name] = f">{args.timeout:.0f} s"
    print(f"{n:>7} | {min(ks):15.3f} ms | {e2e:19.1f} ms | {ref['simple_cuda_ref']:>17} | {ref['rotated_cuda_ref']:>17}", flush=True)

# the configuration the authors published (omp_smithW-v1-refinedOrig.cpp -DSKIP_BACKTRACK=1: no maxPos, no critical
# section) on all host cores, and the plain reference with 1 thread (its per-cell `omp critical` makes threads a loss)
import os
print(f"\nCPU comparators on this host ({os.cpu_count()} cpus): seconds of 'Elapsed time for scoring matrix computation'")
for n in (2048, 8192):
    row = [f"{n:>7}"]
    for name, env in (("v1_skipbt_ref", {}), ("omp_smithW_ref", {"OMP_NUM_THREADS": "1"})):
        exe = ROOT / "oracle" / "_ref" / name
        if not exe.exists():
            row.append(f"{name}: not built"); continue
        try:
            out = subprocess.run([str(exe), str(n), str(n)], capture_output=True, text=True, timeout=args.timeout, env={**os.environ, **env})
            mm = re.search(r"scoring matrix computation:\s*([0-9.]+)", out.stdout)
            sec = float(mm.group(1)) if mm else float("nan")
            row.append(f"{name}{' (1 thread)' if env else ' (all cores)'}: {sec:.3f} s = {n * n / sec / 1e9:.3f} GCUPS")
        except subprocess.TimeoutExpired:
            row.append(f"{name}: >{args.timeout:.0f} s")
    print(" | ".join(row), flush=True)
