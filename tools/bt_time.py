"""Developer tool: time the backtrack kernel alone on the 45000 x 45000 pair (SWB_LIB selects the build)."""
import importlib, sys, torch
sys.path.insert(0, '.')
swb = importlib.import_module("smith-waterman_b200")
cols = rows = 45000
dev = torch.device("cuda:0")
a, b = swb.generate(42, cols, rows)
dH = torch.empty((rows + 1) * (cols + 1), dtype=torch.int32, device=dev); dP = torch.empty_like(dH)
mp = swb.fill(a, cols, b, rows, dH, dP)
keep = dP.clone()
for it in range(3):
    dP.copy_(keep); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); n = swb.backtrack(dP, cols + 1, mp); e1.record(); torch.cuda.synchronize()
    print("backtrack", n, "cells", e0.elapsed_time(e1), "ms")
