import importlib, sys, torch
sys.path.insert(0, '.')
swb = importlib.import_module("smith-waterman_b200")
dev = torch.device("cuda:0")
cols = rows = 45000
a, b = swb.generate(42, cols, rows)
a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
timer = swb.KernelTimer(0)
d_pos = torch.zeros(1, dtype=torch.int64, device=dev); d_sc = torch.zeros(1, dtype=torch.int32, device=dev)
ts = []
for r in range(12):
    swb.score_only_async(a_d, cols, b_d, rows, 1, d_pos, d_sc, stream=torch.cuda.current_stream(), timer=timer)
    torch.cuda.synchronize(); ts.append(round(timer.elapsed_ms(), 2))
print("score-only ms:", ts, int(d_pos.item()), int(d_sc.item()))
