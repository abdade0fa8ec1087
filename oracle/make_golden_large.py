#!/usr/bin/env python3
"""Writes tests/golden/large_digests.json: the CPU oracle's results for the BASELINE.json
configurations that are too large to dump (TEST INFRASTRUCTURE ONLY).

    python oracle/make_golden_large.py [--only 100000x100000,45000x45000,...]

For every single-pair configuration the oracle (oracle/sw_oracle.c, itself pinned against the
unmodified reference by tests/test_oracle.py) sweeps the matrix in blocks of 1024 rows and records,
for a fixed sample of row blocks, the position-weighted digests (oracle/digest.py) of H and P per
column chunk (chunk = the column strip of one GPU when the pair is split over 8), plus maxScore,
maxPos (reference tie-break), the backtrack's path length and a digest of the path cells.  P is kept
2-bit packed for the backtrack (2.5 GB at 100000 x 100000).  For the batch configuration (65536
pairs of 256 x 256, pair k seeded 1000+k) every 16th pair is recorded.

Takes ~6 minutes on one core for everything.  The sequences are the reference's generate()
(omp_smithW.c:489-519) for the seed, so the GPU tests regenerate them with swb.generate().
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.swo import Oracle                      # noqa: E402
from oracle.digest import digest_np, path_digest   # noqa: E402

OUT = ROOT / "tests" / "golden" / "large_digests.json"
BLOCK = 1024

# name -> (cols, rows, seed, chunk_cols, sample stride of row blocks)
SINGLE = {
    "45000x45000": (45000, 45000, 42, 5625, 8),
    "100000x100000": (100000, 100000, 42, 12500, 8),
    "1000x2000000": (1000, 2000000, 42, 125, 128),
    "2000000x1000": (2000000, 1000, 42, 250000, 1),
}
BATCH = {"pairs": 65536, "cols": 256, "rows": 256, "seed0": 1000, "stride": 16}


def sampled(k: int, nblocks: int, stride: int) -> bool:
    return k < 2 or k >= nblocks - 2 or k % stride == 0


def single(orc: Oracle, cols: int, rows: int, seed: int, chunk: int, stride: int) -> dict:
    a, b = orc.generate(seed, cols, rows)
    nblocks = (rows + BLOCK - 1) // BLOCK
    w4 = (cols + 1 + 3) // 4
    packed = np.zeros((rows + 1, w4), dtype=np.uint8)
    blocks = {}
    res = {}
    t0 = time.time()
    for k, (i0, i1, Hb, Pb) in enumerate(orc.fill_blocks(a, b, BLOCK)):
        if i0 is None:
            res["maxScore"], res["maxPos"] = int(Hb), int(Pb)
            break
        pad = np.zeros((Pb.shape[0], 4 * w4), dtype=np.uint8)
        pad[:, :cols + 1] = Pb
        packed[i0:i1] = pad[:, 0::4] | (pad[:, 1::4] << 2) | (pad[:, 2::4] << 4) | (pad[:, 3::4] << 6)
        if sampled(k, nblocks, stride):
            hd, pd = [], []
            for c0 in range(1, cols + 1, chunk):
                c1 = min(c0 + chunk, cols + 1)
                hd.append(list(digest_np(Hb[:, c0:c1], i0, c0)))
                pd.append(list(digest_np(Pb[:, c0:c1], i0, c0)))
            assert not Hb[:, 0].any() and not Pb[:, 0].any()
            blocks[str(i0)] = {"H": hd, "P": pd}
        if k % 16 == 0:
            print(f"  {cols}x{rows}: block {k}/{nblocks}  {time.time() - t0:.0f}s", flush=True)
    # backtrack on the packed P (omp_smithW.c:405-420)
    pos = res["maxPos"]
    path = []
    if pos > 0:
        i, j = divmod(pos, cols + 1)
        while True:
            code = (int(packed[i, j >> 2]) >> (2 * (j & 3))) & 3
            if code == 0:
                break
            path.append(i * (cols + 1) + j)
            if code == 3:
                i, j = i - 1, j - 1
            elif code == 1:
                i -= 1
            else:
                j -= 1
        res["path_end"] = i * (cols + 1) + j
    res.update(cols=cols, rows=rows, seed=seed, block_rows=BLOCK, chunk_cols=chunk, path_len=len(path),
               path_digest=path_digest(path), blocks=blocks)
    return res


def batch(orc: Oracle) -> dict:
    out = {"cols": BATCH["cols"], "rows": BATCH["rows"], "seed0": BATCH["seed0"], "pairs": BATCH["pairs"],
           "stride": BATCH["stride"], "sample": {}}
    for k in range(0, BATCH["pairs"], BATCH["stride"]):
        a, b = orc.generate(BATCH["seed0"] + k, BATCH["cols"], BATCH["rows"])
        H, P, mp = orc.fill(a, b, order="rowmajor")
        plen = orc.backtrack(P.copy(), mp)
        out["sample"][str(k)] = [int(H.reshape(-1)[mp]) if mp else 0, mp, plen, *digest_np(H, 0, 0), *digest_np(P, 0, 0)]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    want = set(x for x in args.only.split(",") if x)
    data = json.loads(OUT.read_text()) if OUT.exists() else {"single": {}, "batch": None}
    orc = Oracle()
    for name, cfg in SINGLE.items():
        if want and name not in want:
            continue
        t0 = time.time()
        data["single"][name] = single(orc, *cfg)
        print(f"{name}: maxPos {data['single'][name]['maxPos']} path {data['single'][name]['path_len']}  {time.time() - t0:.0f}s", flush=True)
        OUT.write_text(json.dumps(data))
    if not want or "batch" in want:
        data["batch"] = batch(orc)
        OUT.write_text(json.dumps(data))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
