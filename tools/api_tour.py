#!/usr/bin/env python3
"""Developer tool: a small tour of the C ABI (a quick end-to-end check on a GPU box): single fills of the three geometries,
score-only, a batch, variable-length pairs, backtrack, traceback, packed transfer -- each checked against the oracle."""
import importlib, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
swb = importlib.import_module("smith-waterman_b200")
from oracle.swo import Oracle
oracle = Oracle()
dev = torch.device("cuda:0")
ACGT = np.frombuffer(b"ACGT", np.uint8)
rng = np.random.default_rng(3)
for (m, n) in [(300, 200), (200, 300), (1100, 130), (97, 5)]:
    a, b = rng.choice(ACGT, m), rng.choice(ACGT, n)
    Ho, Po, mpo = oracle.fill(a, b)
    dH = torch.empty((n + 1) * (m + 1), dtype=torch.int32, device=dev); dP = torch.empty_like(dH)
    assert swb.fill(bytes(a), m, bytes(b), n, dH, dP) == mpo
    assert (dH.view(n + 1, m + 1).cpu().numpy() == Ho).all() and (dP.view(n + 1, m + 1).cpu().numpy() == Po).all()
    assert swb.score_only(bytes(a), bytes(b)) == oracle.score_only(a, b)
    leno = oracle.backtrack(Po, mpo)
    assert swb.backtrack(dP, m + 1, mpo) == leno and (dP.view(n + 1, m + 1).cpu().numpy() == Po).all()
    H = np.empty((n + 1, m + 1), np.int32); P = np.empty_like(H)
    nb = swb.d2h_packed_scratch_bytes(n + 1, m + 1)
    d_s = torch.empty(nb, dtype=torch.uint8, device=dev); h_s = torch.empty(nb, dtype=torch.uint8).pin_memory()
    swb.d2h_packed(dH, dP, m + 1, n + 1, m + 1, H, P, m + 1, d_s, h_s, threads=2, stream=torch.cuda.current_stream())
    assert (H == Ho).all() and (P == Po).all()
    print("single", m, n, "ok", flush=True)
m, n, npairs = 130, 70, 9
A, B = rng.choice(ACGT, (npairs, m)), rng.choice(ACGT, (npairs, n))
stride = ((n + 1) * (m + 1) + 3) // 4 * 4
dH = torch.empty(npairs * stride, dtype=torch.int32, device=dev); dP = torch.empty_like(dH)
pos = torch.zeros(npairs, dtype=torch.int64, device=dev); sc = torch.zeros(npairs, dtype=torch.int32, device=dev)
swb.fill_batch_async(torch.from_numpy(A).to(dev), m, torch.from_numpy(B).to(dev), n, npairs, dH, dP, m + 1, stride, pos, sc, stream=torch.cuda.current_stream())
torch.cuda.synchronize()
for k in range(npairs):
    Ho, Po, mpo = oracle.fill(A[k], B[k])
    assert (dH[k * stride:k * stride + (n + 1) * (m + 1)].view(n + 1, m + 1).cpu().numpy() == Ho).all() and int(pos[k]) == mpo
print("batch ok", flush=True)
