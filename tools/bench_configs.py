#!/usr/bin/env python3
"""Developer tool: kernel-timed GCUPS of the BASELINE.json configurations that fit one GPU.
  python tools/bench_configs.py [--configs square,skew,skewT,batch,score,score_batch,big]"""
import argparse, importlib, sys, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
swb = importlib.import_module("smith-waterman_b200")
dev = torch.device("cuda:0")
PEAK = 6522.7
try:
    PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    pass


def timeit(fn, timer, reps=5):
    ts = []
    for r in range(reps + 2):
        fn(); torch.cuda.synchronize()
        if r >= 2: ts.append(timer.elapsed_ms())
    return min(ts), sorted(ts)[len(ts) // 2]


def single(cols, rows, label, store=True, pitch=None):
    a, b = swb.generate(42, cols, rows)
    a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
    timer = swb.KernelTimer(0)
    d_pos = torch.zeros(1, dtype=torch.int64, device=dev); d_sc = torch.zeros(1, dtype=torch.int32, device=dev)
    if store:
        pitch = pitch or cols + 1
        cells = (rows + 1) * pitch
        dH = torch.empty(cells, dtype=torch.int32, device=dev); dP = torch.empty(cells, dtype=torch.int32, device=dev)
        fn = lambda: swb.fill_async(a_d, cols, b_d, rows, dH, dP, pitch, d_pos, d_sc, stream=torch.cuda.current_stream(), timer=timer)
    else:
        fn = lambda: swb.score_only_async(a_d, cols, b_d, rows, 1, d_pos, d_sc, stream=torch.cuda.current_stream(), timer=timer)
    best, med = timeit(fn, timer)
    g = cols * rows / best / 1e6
    line = f"{label:34s} kernel best {best:9.3f} ms med {med:9.3f} ms  {g:8.1f} GCUPS"
    if store: line += f"  {8.0 * (rows + 1) * (cols + 1) / best / 1e6:8.1f} GB/s = {100 * 8.0 * (rows + 1) * (cols + 1) / best / 1e6 / PEAK:5.1f}% of measured HBM peak"
    print(line + f"  maxPos {int(d_pos.item())} score {int(d_sc.item())}", flush=True)


def batch(m, n, npairs, label, store=True):
    rng = np.random.default_rng(1)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    A = torch.from_numpy(rng.choice(acgt, (npairs, m))).to(dev); B = torch.from_numpy(rng.choice(acgt, (npairs, n))).to(dev)
    timer = swb.KernelTimer(0)
    d_pos = torch.zeros(npairs, dtype=torch.int64, device=dev); d_sc = torch.zeros(npairs, dtype=torch.int32, device=dev)
    pitch = m + 1; stride = ((n + 1) * pitch + 3) // 4 * 4
    if store:
        dH = torch.empty(npairs * stride, dtype=torch.int32, device=dev); dP = torch.empty(npairs * stride, dtype=torch.int32, device=dev)
        fn = lambda: swb.fill_batch_async(A, m, B, n, npairs, dH, dP, pitch, stride, d_pos, d_sc, stream=torch.cuda.current_stream(), timer=timer)
    else:
        fn = lambda: swb.score_only_async(A, m, B, n, npairs, d_pos, d_sc, stream=torch.cuda.current_stream(), timer=timer)
    best, med = timeit(fn, timer, reps=3)
    g = m * n * npairs / best / 1e6
    line = f"{label:34s} kernel best {best:9.3f} ms med {med:9.3f} ms  {g:8.1f} GCUPS"
    if store: line += f"  {8.0 * npairs * (n + 1) * pitch / best / 1e6:8.1f} GB/s = {100 * 8.0 * npairs * (n + 1) * pitch / best / 1e6 / PEAK:5.1f}% of measured HBM peak"
    print(line, flush=True)


ap = argparse.ArgumentParser(); ap.add_argument("--configs", default="square,skew,skewT,batch,score,score_batch")
args = ap.parse_args()
for c in args.configs.split(","):
    if c == "square": single(45000, 45000, "45000x45000 full fill")
    if c == "skew": single(1000, 2000000, "1000 cols x 2000000 rows full fill")
    if c == "skewT": single(2000000, 1000, "2000000 cols x 1000 rows full fill")
    if c == "batch": batch(256, 256, 65536, "65536 x 256x256 full fill")
    if c == "score": single(45000, 45000, "45000x45000 score only", store=False)
    if c == "score_batch": batch(256, 256, 65536, "65536 x 256x256 score only", store=False)
    if c == "pitch":
        for pt in (45001, 45056, 45088, 45312, 46081, 49152, 49153):
            single(45000, 45000, f"45000x45000 pitch {pt}", pitch=pt)
        for sz in (28416, 42624, 44928, 56832, 71040):   # 296, 444, 468, 592, 740 strips of 96 rows
            single(sz, sz, f"{sz}x{sz} ({sz // 96} strips)")
    if c == "big": single(100000, 100000, "100000x100000 full fill (80 GB)")
    torch.cuda.empty_cache()
