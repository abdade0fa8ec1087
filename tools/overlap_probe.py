#!/usr/bin/env python3
"""Developer tool: N fills of 45000x45000 on two streams (no backtrack), different ways of staggering the starts."""
import importlib, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
swb = importlib.import_module("smith-waterman_b200")
dev = torch.device("cuda:0")
n = m = 45000
a, b = swb.generate(42, m, n)
a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
cells = (n + 1) * (m + 1)
NS = 4
sets = [(torch.empty(cells, dtype=torch.int32, device=dev), torch.empty(cells, dtype=torch.int32, device=dev),
         torch.zeros(1, dtype=torch.int64, device=dev)) for _ in range(NS)]
fs = [torch.cuda.Stream(device=dev) for _ in range(2)]


def run(nfills, sleep_us, timers=None):
    for k in range(nfills):
        H_, P_, sc_ = sets[k % NS]
        swb.fill_async(a_d, m, b_d, n, H_, P_, m + 1, sc_, None, device=0, stream=fs[k % 2], timer=timers[k] if timers else None)
        if sleep_us and k == 0:
            t = time.perf_counter()
            while (time.perf_counter() - t) * 1e6 < sleep_us:
                pass


for nfills in (1, 2, 3, 4):
    for sleep_us in (0, 1000, 2000, 3000):
        if nfills == 1 and sleep_us:
            continue
        best = None
        for rep in range(3):
            timers = [swb.KernelTimer(0) for _ in range(nfills)]
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record(fs[0]); fs[1].wait_event(t0)
            run(nfills, sleep_us, timers)
            e = torch.cuda.Event(); e.record(fs[1]); fs[0].wait_event(e)
            t1.record(fs[0]); torch.cuda.synchronize()
            ms = t0.elapsed_time(t1)
            if best is None or ms < best[0]:
                best = (ms, [round(t.elapsed_ms(), 2) for t in timers])
        print(f"fills {nfills} host stagger {sleep_us:5d} us: total {best[0]:7.3f} ms  per-kernel {best[1]}", flush=True)
