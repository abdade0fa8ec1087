"""Developer tool: SM clock seen by each strip = clock64 ticks / globaltimer ns between its gate and its end
(build with -DSWB_X_CLKTRACE)."""
import importlib, sys, torch, numpy as np
sys.path.insert(0, '.')
swb = importlib.import_module("smith-waterman_b200")
cols = rows = 45000
dev = torch.device("cuda:0")
a, b = swb.generate(42, cols, rows)
a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
dH = torch.empty((rows + 1) * (cols + 1), dtype=torch.int32, device=dev); dP = torch.empty_like(dH)
strips = (rows + 95) // 96
for it in range(3):
    tr = torch.zeros(strips * 8, dtype=torch.int64, device=dev)
    swb.fill_async(a_d, cols, b_d, rows, dH, dP, cols + 1, None, None, warps_per_band=2, trace=tr)
    torch.cuda.synchronize()
t = tr.view(strips, 8).cpu().numpy().astype("float64")
ghz = t[:, 6] / (t[:, 4] - t[:, 1])
for lo in range(0, strips, strips // 8):
    print("strips %4d-%4d: SM clock %.3f GHz (min %.3f max %.3f), strip time %.1f us" % (lo, min(strips, lo + strips // 8) - 1, ghz[lo:lo + strips // 8].mean(),
          ghz[lo:lo + strips // 8].min(), ghz[lo:lo + strips // 8].max(), (t[lo:lo + strips // 8, 4] - t[lo:lo + strips // 8, 1]).mean() / 1000))
