"""Developer tool: strip timeline of the score-only kernel for a few launches (fast and slow ones)."""
import importlib, sys, torch, numpy as np
sys.path.insert(0, '.')
swb = importlib.import_module("smith-waterman_b200")
dev = torch.device("cuda:0")
cols = rows = 45000
a, b = swb.generate(42, cols, rows)
a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
timer = swb.KernelTimer(0)
strips = (rows + 63) // 64
for r in range(8):
    tr = torch.zeros(strips * 8, dtype=torch.int64, device=dev)
    swb.score_only_async(a_d, cols, b_d, rows, 1, None, None, stream=torch.cuda.current_stream(), timer=timer, trace=tr)
    torch.cuda.synchronize()
    t = tr.view(strips, 8).cpu().numpy().astype("float64"); t = (t - t[:, 0].min()) / 1000.0
    lag = np.diff(t[:, 1])
    dur = t[:, 4] - t[:, 1]
    print("launch %d: %.2f ms | last gate %.0f us, strip duration mean %.0f us (min %.0f max %.0f) | lag mean %.2f us median %.2f p99 %.2f max %.1f | strips with lag > 20 us: %s" % (
        r, timer.elapsed_ms(), t[-1, 1], dur.mean(), dur.min(), dur.max(), lag.mean(), np.median(lag), np.percentile(lag, 99), lag.max(),
        list(np.nonzero(lag > 20)[0][:12])))
