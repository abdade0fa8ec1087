"""Developer tool: strip timeline of the score-only kernel for a few launches (fast and slow ones)."""
import importlib, os, sys, torch, numpy as np
os.environ.setdefault("SWB_LIB", "build/libswb200_trace.so")   # make -C smith-waterman_b200 trace
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
        list(np.nonzero(lag > 20)[0][:6])))
    inner, cross = lag[0::2], lag[1::2]
    print("   inner lag median %.2f p90 %.2f | cross lag median %.2f p90 %.2f | duration even strips median %.0f, odd %.0f | first 32 steps (g4-gate) median %.2f us, next 32 %.2f us, steady ns/step %.1f" % (
        np.median(inner), np.percentile(inner, 90), np.median(cross), np.percentile(cross, 90), np.median(dur[0::2]), np.median(dur[1::2]),
        np.median(t[:, 2] - t[:, 1]), np.median(t[:, 3] - t[:, 2]), np.median((t[:, 4] - t[:, 3]) * 1000 / (cols // 4 - 32))))
    raw = tr.view(strips, 8).cpu().numpy()
    slow = np.nonzero(dur > 3 * dur.min())[0]
    print("   strips slower than 3x the fastest: %d; first ones (strip, SM, warp slot): %s" % (len(slow), [(int(s_), int(raw[s_, 7] & 0xffffffff), int(raw[s_, 7] >> 32)) for s_ in slow[:4]]),
          "| warp slots of normal strips:", sorted(set(int(x >> 32) for x in raw[dur < 1.5 * dur.min(), 7]))[:12])
    for s_ in slow[:3]:
        print("      strip %d: enter %.0f gate %.0f g4 %.0f g8 %.0f end %.0f us | prev strip: enter %.0f gate %.0f g4 %.0f g8 %.0f end %.0f" % (
            s_, *t[s_, :5], *t[s_ - 1, :5]))
    k = int(np.argmax(lag > 20)) if (lag > 20).any() else 0
    print("   around the first slow strip %d: gate %s" % (k, " ".join("%.0f" % x for x in t[max(0, k - 3):k + 6, 1])), "| end", " ".join("%.0f" % x for x in t[max(0, k - 3):k + 6, 4]))
