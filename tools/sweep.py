#!/usr/bin/env python3
"""Developer tool: times the fill kernel (CUDA events around the kernel alone) over
shapes and warps-per-band settings.  Not part of the product or the bench contract.

  python tools/sweep.py --shapes 45000x45000,2048x2048 --wpc 2,4,8 --reps 5
"""
import argparse
import importlib
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
swb = importlib.import_module("smith-waterman_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="45000x45000")
    ap.add_argument("--wpc", default="4")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    timer = swb.KernelTimer(0)
    for shp in args.shapes.split(","):
        cols, rows = (int(x) for x in shp.split("x"))
        a, b = swb.generate(args.seed, cols, rows)
        a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev)
        b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
        cells = (rows + 1) * (cols + 1)
        dH = torch.empty(cells, dtype=torch.int32, device=dev)
        dP = torch.empty(cells, dtype=torch.int32, device=dev)
        d_pos = torch.zeros(1, dtype=torch.int64, device=dev)
        for wpc in (int(x) for x in args.wpc.split(",")):
            times, tot = [], []
            for r in range(args.reps + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                swb.fill_async(a_d, cols, b_d, rows, dH, dP, cols + 1, d_pos, None, device=0,
                               stream=torch.cuda.current_stream(), warps_per_band=wpc, timer=timer)
                e1.record()
                torch.cuda.synchronize()
                if r >= 2:
                    times.append(timer.elapsed_ms()); tot.append(e0.elapsed_time(e1))
            best, med = min(times), sorted(times)[len(times) // 2]
            gcups = cols * rows / (best * 1e-3) / 1e9
            gbs = 8.0 * cells / (best * 1e-3) / 1e9
            print(f"{cols}x{rows} wpc={wpc:2d} kernel best {best:9.3f} ms med {med:9.3f} ms  total(best) {min(tot):9.3f} ms"
                  f"  {gcups:8.1f} GCUPS  {gbs:8.1f} GB/s written  maxPos {int(d_pos.item())}", flush=True)
        del dH, dP
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
