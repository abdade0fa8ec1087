"""Developer tool: per-group clock counters of the score-only kernel (build with -DSWB_X_GROUPTRACE), slow strips first."""
import importlib, sys, torch, numpy as np
sys.path.insert(0, '.')
swb = importlib.import_module("smith-waterman_b200")
dev = torch.device("cuda:0")
cols = rows = 45000
a, b = swb.generate(42, cols, rows)
a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
strips = (rows + 63) // 64
for r in range(6):
    tr = torch.zeros(strips * 8, dtype=torch.int64, device=dev)
    swb.score_only_async(a_d, cols, b_d, rows, 1, None, None, stream=torch.cuda.current_stream(), trace=tr)
    torch.cuda.synchronize()
    t = tr.view(strips, 8).cpu().numpy().astype(object)
    n = np.maximum(t[:, 5].astype(float), 1)
    per = (t[:, 2].astype(float) + t[:, 3].astype(float) + t[:, 4].astype(float)) / n
    slow = np.nonzero(per > 2.5 * np.median(per[:8]))[0]
    print("launch %d: median clk/group %.0f; slow strips %d; first: %s" % (r, np.median(per), len(slow), list(slow[:4])))
    for s in list(slow[:3]) + [0, 1]:
        x = t[s]; k = max(float(x[5]), 1)
        drain = (int(x[7]) >> 32) * 16; cons = (int(x[7]) & 0xffffffff) * 16
        print("   strip %4d: pre %.0f (writers' ring %.0f, consumer's ring %.0f)  steps %.0f (%.0f/step)  post %.0f | head %.0f/step" % (
            s, x[2] / k, drain / k, cons / k, x[3] / k, x[3] / k / 8, x[4] / k, x[6] / 32))
