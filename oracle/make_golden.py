#!/usr/bin/env python3
"""Generates tests/golden/ from the UNMODIFIED reference (oracle/_ref/libswref.so).

Run it where /root/reference exists (`make -C oracle ref` first).  The fixtures
travel with the repo; /root/reference does not exist on the GPU box.

  tests/golden/ref_small.npz   full dumps (a, b, H, post-backtrack P, maxPos) of
                               the built-in case and small random cases
  tests/golden/ref_hashes.json FNV-1a digests of H and post-backtrack P, maxPos
                               and path length for larger cases (2048^2 = BASELINE
                               configs[0], skewed shapes, tie-heavy 256^2 runs)

Sequences come from the reference's own generate() with time() pinned to the
seed (omp_smithW.c:489-519), OMP_NUM_THREADS=1 (SURVEY.md 8c determinism note).
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.swo import Oracle, Reference  # noqa: E402

SMALL = [  # (cols, rows, seed); cols<=0 = built-in
    (0, 0, 0),
    (1, 1, 1), (1, 1, 2), (2, 2, 2), (3, 5, 3), (1, 50, 3), (50, 1, 4), (7, 7, 11),
    (31, 33, 12), (32, 32, 13), (33, 31, 14), (64, 64, 7), (65, 63, 15), (37, 301, 5),
    (301, 37, 6), (5, 1000, 8), (1000, 5, 9), (128, 129, 16), (130, 127, 17), (129, 128, 18),
]
HASHED = [
    (2048, 2048, 1), (2048, 2048, 42), (2048, 2048, 20260101),   # BASELINE configs[0]
    (256, 256, 1000), (256, 256, 1001), (256, 256, 1002), (256, 256, 1003),  # batch-config pairs
    (255, 257, 10), (1000, 1000, 42), (1000, 3000, 42), (3000, 1000, 42),
    (45, 4000, 21), (4000, 45, 22), (1021, 2053, 23), (2053, 1021, 24), (4099, 515, 25),
]


def main() -> None:
    ref, orc = Reference(), Oracle()
    out = ROOT / "tests" / "golden"
    out.mkdir(parents=True, exist_ok=True)

    small = {}
    for k, (c, r, s) in enumerate(SMALL):
        res = ref.run(c, r, s, threads=1)
        tag = f"c{k:02d}"
        small[f"{tag}_meta"] = np.array([c, r, s, res["maxPos"], res["path_len"]], dtype=np.int64)
        small[f"{tag}_a"] = res["a"]
        small[f"{tag}_b"] = res["b"]
        small[f"{tag}_H"] = res["H"]
        small[f"{tag}_Pbt"] = res["P_bt"]
    np.savez_compressed(out / "ref_small.npz", **small)

    hashed = []
    for (c, r, s) in HASHED:
        res = ref.run(c, r, s, threads=1)
        hashed.append(dict(cols=c, rows=r, seed=s, maxPos=res["maxPos"], path_len=res["path_len"],
                           maxScore=int(res["H"].max()),
                           H_fnv=f"{orc.fnv(res['H']):016x}", Pbt_fnv=f"{orc.fnv(res['P_bt']):016x}",
                           a_head=bytes(res["a"][:16]).decode(), b_head=bytes(res["b"][:16]).decode()))
        print(hashed[-1])
    (out / "ref_hashes.json").write_text(json.dumps(
        dict(source="unmodified /root/reference/omp_smithW.c via oracle/ref_shim.cpp, OMP_NUM_THREADS=1",
             scoring=[3, -3, -2], cases=hashed), indent=1) + "\n")


if __name__ == "__main__":
    main()
