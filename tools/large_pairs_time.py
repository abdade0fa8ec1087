#!/usr/bin/env python3
"""Developer tool: time swb_fill_pairs_async on k large pairs (overlap on two internal streams)."""
import importlib, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
swb = importlib.import_module("smith-waterman_b200")
dev = torch.device("cuda:0")
c = r = 45000
a, b = swb.generate(42, c, r)
size = ((c + 1) * (r + 1) + 3) // 4 * 4
NP = 6
A = torch.frombuffer(bytearray(a * NP), dtype=torch.uint8).to(dev); B = torch.frombuffer(bytearray(b * NP), dtype=torch.uint8).to(dev)
dH = torch.empty(NP * size, dtype=torch.int32, device=dev); dP = torch.empty_like(dH)
pos = torch.zeros(NP, dtype=torch.int64, device=dev)
for user in ("default", "side"):
    st = torch.cuda.current_stream() if user == "default" else torch.cuda.Stream(device=dev)
    for npairs in (1, 2, 3, 4, 6):
        offs = lambda step: [k * step for k in range(npairs)]
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(st)
            swb.fill_pairs_async(A, offs(c), [c] * npairs, B, offs(r), [r] * npairs, offs(size), dH, dP, pos, None, stream=st)
            e1.record(st); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"user stream {user:8s} pairs {npairs}: {best:8.3f} ms total, {best / npairs:6.3f} ms per pair", flush=True)
