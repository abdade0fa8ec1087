"""Developer tool: per-strip step times for an arbitrary shape (few strips)."""
import importlib, sys, torch, numpy as np
sys.path.insert(0, '.')
swb = importlib.import_module("smith-waterman_b200")
cols, rows, wpc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
SR = 64
dev = torch.device("cuda:0")
a, b = swb.generate(42, cols, rows)
a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
dH = torch.empty((rows + 1) * (cols + 1), dtype=torch.int32, device=dev); dP = torch.empty_like(dH)
strips = (rows + SR - 1) // SR
for it in range(3):
    tr = torch.zeros(strips * 8, dtype=torch.int64, device=dev)
    swb.fill_async(a_d, cols, b_d, rows, dH, dP, cols + 1, None, None, warps_per_band=wpc, trace=tr)
    torch.cuda.synchronize()
t = tr.view(strips, 8).cpu().numpy().astype("float64"); t = (t - t[:, 0].min()) / 1000.0
nsteps = cols // 4 + 32
for s in range(min(strips, 8)):
    print(f"strip {s}: gate {t[s,1]:8.1f} us  first32 {1000*(t[s,2]-t[s,1])/32:6.1f} ns/step  next32 {1000*(t[s,3]-t[s,2])/32:6.1f}  rest {1000*(t[s,4]-t[s,3])/(nsteps-64):6.1f} ns/step  (x1.962 = clk)")
