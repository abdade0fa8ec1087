#!/usr/bin/env python3
"""Developer tool: throughput of consecutive 45000x45000 steps (fill + maxPos + backtrack) when the fills alternate
between TWO streams, so that the first strips of step k+1 run on the SMs the ramp-down of step k leaves idle.
  python tools/overlap_fills.py [--sets 3] [--steps 12] [--streams 2]"""
import argparse, importlib, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
swb = importlib.import_module("smith-waterman_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--sets", type=int, default=3); ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--streams", type=int, default=2); ap.add_argument("--size", type=int, default=45000)
args = ap.parse_args()
dev = torch.device("cuda:0")
n = m = args.size
a, b = swb.generate(42, m, n)
a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
cells = (n + 1) * (m + 1)
sets = [(torch.empty(cells, dtype=torch.int32, device=dev), torch.empty(cells, dtype=torch.int32, device=dev),
         torch.zeros(2, dtype=torch.int64, device=dev)) for _ in range(args.sets)]
fs = [torch.cuda.Stream(device=dev) for _ in range(args.streams)]
s_bt = torch.cuda.Stream(device=dev, priority=-1)


def run(nsteps):
    e_fill = [torch.cuda.Event() for _ in range(nsteps)]; e_bt = [torch.cuda.Event() for _ in range(nsteps)]
    for k in range(nsteps):
        H_, P_, sc_ = sets[k % args.sets]
        s = fs[k % args.streams]
        if k >= args.sets:
            s.wait_event(e_bt[k - args.sets])
        swb.fill_async(a_d, m, b_d, n, H_, P_, m + 1, sc_[0:1], None, device=0, stream=s)
        e_fill[k].record(s)
        s_bt.wait_event(e_fill[k])
        swb.backtrack_async(P_, m + 1, d_maxPos=sc_[0:1], d_pathLen=sc_[1:2], device=0, stream=s_bt)
        e_bt[k].record(s_bt)
    for k in range(max(0, nsteps - args.sets), nsteps):
        for s in fs:
            s.wait_event(e_bt[k])


run(4); torch.cuda.synchronize()
for rep in range(3):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record(fs[0])
    for s in fs[1:]:
        s.wait_event(t0)
    run(args.steps)
    t1.record(fs[0]); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    print(f"streams {args.streams} sets {args.sets}: {ms:.3f} ms per step = {m * n / ms / 1e6:.1f} GCUPS   results {[tuple(int(x) for x in sc.tolist()) for _, _, sc in sets[:2]]}", flush=True)
