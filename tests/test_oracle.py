"""Pins the CPU oracle (oracle/sw_oracle.c) to the reference.

1. the reference's own known-answer test (omp_smithW.c:147-164,230-234 and the
   stronger asserts of omp_smithW-v1-refinedOrig.cpp:231-237),
2. dumps of the unmodified reference committed under tests/golden/
   (oracle/make_golden.py), full matrices for small cases and digests for
   larger ones incl. BASELINE configs[0] (2048x2048),
3. internal consistency: wavefront order == row-major order + explicit
   tie-break; score-only == full fill.
"""
import json

import numpy as np
import pytest

BUILTIN_A = b"TGTTACGG"      # omp_smithW.c:157-164 (columns)
BUILTIN_B = b"GGTTGACTA"     # omp_smithW.c:147-155 (rows)
BUILTIN_H = np.array([
    [0, 0, 0, 0, 0, 0, 0, 0, 0],
    [0, 0, 3, 1, 0, 0, 0, 3, 3],
    [0, 0, 3, 1, 0, 0, 0, 3, 6],
    [0, 3, 1, 6, 4, 2, 0, 1, 4],
    [0, 3, 1, 4, 9, 7, 5, 3, 2],
    [0, 1, 6, 4, 7, 6, 4, 8, 6],
    [0, 0, 4, 3, 5, 10, 8, 6, 5],
    [0, 0, 2, 1, 3, 8, 13, 11, 9],
    [0, 3, 1, 5, 4, 6, 11, 10, 8],
    [0, 1, 0, 3, 2, 7, 9, 8, 7]], dtype=np.int32)
BUILTIN_PBT = np.array([
    [0, 0, 0, 0, 0, 0, 0, 0, 0],
    [0, 0, 3, 2, 0, 0, 0, 3, 3],
    [0, 0, -3, 2, 0, 0, 0, 3, 3],
    [0, 3, 1, -3, 3, 2, 0, 1, 1],
    [0, 3, 2, 3, -3, 2, 2, 2, 1],
    [0, 1, 3, 2, -1, 3, 3, 3, 3],
    [0, 0, 1, 3, 1, -3, 2, 1, 3],
    [0, 0, 1, 3, 1, 1, -3, 2, 2],
    [0, 3, 2, 3, 3, 1, 1, 3, 3],
    [0, 1, 0, 1, 3, 3, 1, 3, 3]], dtype=np.int32)


@pytest.mark.parametrize("order", ["wavefront", "rowmajor"])
def test_builtin_known_answer(oracle, order):
    H, P, maxPos = oracle.fill(BUILTIN_A, BUILTIN_B, order=order)
    assert H[9, 8] == 7                      # omp_smithW.c:233  H[n*m-1]==7
    assert maxPos == 69                      # v1:233
    assert H.reshape(-1)[maxPos] == 13       # v1:234
    assert (H == BUILTIN_H).all()
    assert (P == np.abs(BUILTIN_PBT)).all()
    n = oracle.backtrack(P, maxPos)
    assert n == 6 and (P == BUILTIN_PBT).all()
    assert sorted(np.flatnonzero(P.reshape(-1) < 0)) == [20, 30, 40, 49, 59, 69]


def _small_cases(golden_dir):
    z = np.load(golden_dir / "ref_small.npz")
    tags = sorted({k.split("_")[0] for k in z.files})
    for t in tags:
        yield t, z[f"{t}_meta"], z[f"{t}_a"], z[f"{t}_b"], z[f"{t}_H"], z[f"{t}_Pbt"]


def test_small_dumps_of_the_reference(oracle, golden_dir):
    n_cases = 0
    for tag, meta, a, b, H_ref, Pbt_ref in _small_cases(golden_dir):
        cols, rows, seed, maxPos_ref, plen_ref = (int(x) for x in meta)
        if cols > 0:   # the generator restatement must reproduce the sequences
            ga, gb = oracle.generate(seed, cols, rows)
            assert (ga == a).all() and (gb == b).all(), tag
        for order in ("wavefront", "rowmajor"):
            H, P, maxPos = oracle.fill(a, b, order=order)
            assert (H == H_ref).all(), (tag, order)
            assert (P == np.abs(Pbt_ref)).all(), (tag, order)
            plen = oracle.backtrack(P, maxPos)
            assert (P == Pbt_ref).all(), (tag, order)
            assert plen == plen_ref, (tag, order)
            if plen_ref:
                assert maxPos == maxPos_ref, (tag, order)
            else:
                assert maxPos == 0
        n_cases += 1
    assert n_cases >= 20


def test_hashed_dumps_of_the_reference(oracle, golden_dir):
    meta = json.loads((golden_dir / "ref_hashes.json").read_text())
    assert meta["scoring"] == [3, -3, -2]
    for c in meta["cases"]:
        a, b = oracle.generate(c["seed"], c["cols"], c["rows"])
        assert bytes(a[:16]).decode() == c["a_head"] and bytes(b[:16]).decode() == c["b_head"]
        H, P, maxPos = oracle.fill(a, b)
        assert maxPos == c["maxPos"], c
        assert int(H.max()) == c["maxScore"]
        assert f"{oracle.fnv(H):016x}" == c["H_fnv"], c
        plen = oracle.backtrack(P, maxPos)
        assert plen == c["path_len"]
        assert f"{oracle.fnv(P):016x}" == c["Pbt_fnv"], c


@pytest.mark.parametrize("seed", range(40))
def test_wavefront_equals_rowmajor_tie_heavy(oracle, seed):
    # 256x256 has a multi-cell global maximum in ~30% of the runs
    rng = np.random.default_rng(seed)
    m, n = (256, 256) if seed % 2 else (int(rng.integers(1, 300)), int(rng.integers(1, 300)))
    a = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), m)
    b = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n)
    H1, P1, mp1 = oracle.fill(a, b, order="wavefront")
    H2, P2, mp2 = oracle.fill(a, b, order="rowmajor")
    assert (H1 == H2).all() and (P1 == P2).all() and mp1 == mp2
    ms, mp3 = oracle.score_only(a, b)
    assert ms == H1.max() and mp3 == mp1


def test_other_scoring_and_degenerate_inputs(oracle):
    a = np.frombuffer(b"AAAAAAAA", dtype=np.uint8)
    b = np.frombuffer(b"CCCC", dtype=np.uint8)
    H, P, mp = oracle.fill(a, b)
    assert H.max() == 0 and P.max() == 0 and mp == 0      # no positive score -> maxPos 0
    assert oracle.backtrack(P, mp) == 0
    H, P, mp = oracle.fill(a, a, scoring=(5, -3, -4))     # upstream's original scores (omp_smithW_orig.c:65-67)
    assert H[8, 8] == 40 and mp == 9 * 8 + 8
    assert oracle.backtrack(P, mp) == 8


def test_diag_enumeration_helpers(oracle):
    # nElement sums to the number of interior cells for skewed and square shapes
    for (m, n) in [(8, 9), (9, 8), (1, 7), (7, 1), (33, 33)]:
        M, N = m + 1, n + 1
        total = sum(oracle.lib.swo_nelement(i, M, N) for i in range(1, M + N - 2))
        assert total == m * n
