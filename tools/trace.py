#!/usr/bin/env python3
"""Developer tool: per-strip globaltimer stamps of the fill kernel (swb_tuning.trace).
  python tools/trace.py --shape 8192x8192 --wpc 2"""
import argparse, importlib, os, sys
os.environ.setdefault("SWB_LIB", "build/libswb200_trace.so")   # make -C smith-waterman_b200 trace
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
swb = importlib.import_module("smith-waterman_b200")
ap = argparse.ArgumentParser(); ap.add_argument("--shape", default="8192x8192"); ap.add_argument("--wpc", type=int, default=2)
args = ap.parse_args()
cols, rows = (int(x) for x in args.shape.split("x"))
dev = torch.device("cuda:0")
a, b = swb.generate(42, cols, rows)
a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
dH = torch.empty((rows + 1) * (cols + 1), dtype=torch.int32, device=dev); dP = torch.empty_like(dH)
strips = (rows + 95) // 96          # single pairs run the 96-row geometry (three rows per lane)
for it in range(2):
    tr = torch.zeros(strips * 8, dtype=torch.int64, device=dev)
    swb.fill_async(a_d, cols, b_d, rows, dH, dP, cols + 1, None, None, warps_per_band=args.wpc, trace=tr)
    torch.cuda.synchronize()
t = tr.view(strips, 8).cpu().numpy().astype("float64")
t0 = t[:, 0].min()
t = (t - t0) / 1000.0
names = ["enter", "gate", "g4", "g8", "end", "w0_end", "w1_end"]
print("strip " + " ".join(f"{n:>9s}" for n in names) + "   (us since first entry)")
for s in list(range(0, min(strips, 12))) + list(range(strips // 2, strips // 2 + 4)) + list(range(strips - 3, strips)):
    print(f"{s:5d} " + " ".join(f"{t[s, k]:9.1f}" for k in range(7)))
import numpy as np
print("mean gate->gate lag between consecutive strips (us):", np.diff(t[:, 1]).mean())
print("mean g4-gate (first 32 steps) us:", (t[:, 2] - t[:, 1]).mean(), " g8-g4 (next 32 steps):", (t[:, 3] - t[:, 2]).mean(),
      " end-g8 per step (ns):", ((t[:, 4] - t[:, 3]) * 1000 / ( (cols//4 + 32) - 64)).mean())
g = t[:, 1]
lag = np.diff(g)
inner = lag[0::2]      # strip 2k -> 2k+1: same CTA (shared-memory ring)
cross = lag[1::2]      # strip 2k+1 -> 2k+2: next band (global boundary row + loader)
print("lag inside a CTA   : mean %.2f us  median %.2f  p90 %.2f  max %.2f" % (inner.mean(), np.median(inner), np.percentile(inner, 90), inner.max()))
print("lag across bands   : mean %.2f us  median %.2f  p90 %.2f  max %.2f" % (cross.mean(), np.median(cross), np.percentile(cross, 90), cross.max()))
for lo in range(0, strips - 1, max(1, strips // 8)):
    hi = min(strips - 1, lo + max(1, strips // 8))
    print("  strips %4d-%4d: mean lag %.2f us (inner %.2f, cross %.2f)" % (lo, hi, lag[lo:hi].mean(), lag[lo:hi][0::2].mean() if lo % 2 == 0 else lag[lo:hi][1::2].mean(), lag[lo:hi][1::2].mean() if lo % 2 == 0 else lag[lo:hi][0::2].mean()))
