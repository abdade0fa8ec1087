"""Developer tool: quick bit-exact check of the fill against the oracle for a few shapes."""
import importlib, sys
import numpy as np, torch
sys.path.insert(0, '.')
swb = importlib.import_module("smith-waterman_b200")
from oracle.swo import Oracle
o = Oracle()
rng = np.random.default_rng(5)
acgt = np.frombuffer(b"ACGT", np.uint8)
ok = True
for (m, n, wpc) in [(8, 9, 1), (300, 200, 1), (1027, 700, 2), (4100, 1500, 1), (700, 4100, 2), (5000, 3000, 3)]:
    a, b = rng.choice(acgt, m), rng.choice(acgt, n)
    dH = torch.full(((n + 1) * (m + 1),), -9, dtype=torch.int32, device="cuda"); dP = torch.full_like(dH, -9)
    dpos = torch.zeros(1, dtype=torch.int64, device="cuda")
    swb.fill_async(a, m, b, n, dH, dP, m + 1, dpos, None, warps_per_band=wpc); torch.cuda.synchronize()
    H, P, mp = o.fill(a, b)
    good = (dH.view(n + 1, m + 1).cpu().numpy() == H).all() and (dP.view(n + 1, m + 1).cpu().numpy() == P).all() and int(dpos.item()) == mp
    print(m, n, wpc, "OK" if good else "MISMATCH"); ok &= bool(good)
sys.exit(0 if ok else 1)
