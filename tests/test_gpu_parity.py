"""GPU parity tests: the CUDA path (through the C ABI, include/swb200.h) against the
CPU oracle and the committed dumps of the unmodified reference.  Bit-exact: all
arithmetic on this path is int32/int64 (omp_smithW.c:331-388).
"""
import json

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def gpu_fill(swb, a, b, scoring=None, wpc=0, pitch=None, backtrack=False):
    """-> H, P (numpy (n+1, m+1)), maxPos, maxScore[, path_len]"""
    a = np.ascontiguousarray(a, dtype=np.uint8); b = np.ascontiguousarray(b, dtype=np.uint8)
    m, n = len(a), len(b)
    pitch = pitch or m + 1
    dev = torch.device("cuda:0")
    # poison the outputs: the fill must write every element of rows 0..n, cols 0..m
    dH = torch.full(((n + 1) * pitch,), -999, dtype=torch.int32, device=dev)
    dP = torch.full(((n + 1) * pitch,), -999, dtype=torch.int32, device=dev)
    d_pos = torch.zeros(1, dtype=torch.int64, device=dev)
    d_sc = torch.zeros(1, dtype=torch.int32, device=dev)
    swb.fill_async(a, m, b, n, dH, dP, pitch, d_pos, d_sc, scoring=scoring, device=0,
                   stream=torch.cuda.current_stream(), warps_per_band=wpc)
    torch.cuda.synchronize()
    maxPos, maxScore = int(d_pos.item()), int(d_sc.item())
    out = []
    if backtrack:
        out.append(swb.backtrack(dP, pitch, maxPos, device=0, stream=torch.cuda.current_stream()))
    H = dH.view(n + 1, pitch)[:, : m + 1].cpu().numpy()
    P = dP.view(n + 1, pitch)[:, : m + 1].cpu().numpy()
    if pitch > m + 1:      # the padding must stay untouched
        assert int((dH.view(n + 1, pitch)[:, m + 1:] != -999).sum()) == 0
        assert int((dP.view(n + 1, pitch)[:, m + 1:] != -999).sum()) == 0
    return (H, P, maxPos, maxScore, *out)


def test_builtin_known_answer(swb, oracle):
    # omp_smithW.c:147-164,230-234; omp_smithW-v1-refinedOrig.cpp:231-237
    a, b = b"TGTTACGG", b"GGTTGACTA"
    H, P, maxPos, maxScore, plen = gpu_fill(swb, np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8),
                                            backtrack=True)
    Ho, Po, mpo = oracle.fill(a, b, order="wavefront")
    assert H[9, 8] == 7 and maxPos == 69 and maxScore == 13
    assert (H == Ho).all()
    leno = oracle.backtrack(Po, mpo)
    assert plen == leno == 6 and (P == Po).all()
    assert sorted(np.flatnonzero(P.reshape(-1) < 0)) == [20, 30, 40, 49, 59, 69]


def test_small_dumps_of_the_reference(swb, golden_dir):
    z = np.load(golden_dir / "ref_small.npz")
    tags = sorted({k.split("_")[0] for k in z.files})
    for t in tags:
        cols, rows, seed, maxPos_ref, plen_ref = (int(x) for x in z[f"{t}_meta"])
        H, P, maxPos, _, plen = gpu_fill(swb, z[f"{t}_a"], z[f"{t}_b"], backtrack=True)
        assert (H == z[f"{t}_H"]).all(), t
        assert (P == z[f"{t}_Pbt"]).all(), t
        assert plen == plen_ref, t
        assert maxPos == (maxPos_ref if plen_ref else 0), t


def test_hashed_dumps_of_the_reference(swb, oracle, golden_dir):
    meta = json.loads((golden_dir / "ref_hashes.json").read_text())
    for c in meta["cases"]:
        a, b = swb.generate(c["seed"], c["cols"], c["rows"])      # the product's generate() clone
        assert a[:16].decode() == c["a_head"] and b[:16].decode() == c["b_head"]
        H, P, maxPos, maxScore, plen = gpu_fill(swb, np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8),
                                                backtrack=True)
        assert maxPos == c["maxPos"] and maxScore == c["maxScore"] and plen == c["path_len"], c
        assert f"{oracle.fnv(H):016x}" == c["H_fnv"], c
        assert f"{oracle.fnv(P):016x}" == c["Pbt_fnv"], c


# alphabets: 4 letters -> the score-profile instantiation of the fill kernel (at most 8 distinct bytes in b);
# 12 letters -> the character-compare instantiation; 8 letters incl. NUL -> the profile's padding must not match NUL
ALPHABETS = {"dna": ACGT, "wide": np.frombuffer(b"ACDEFGHIKLMN", dtype=np.uint8),
             "nul8": np.array([0, 65, 67, 71, 84, 78, 1, 255], dtype=np.uint8),
             "nul12": np.arange(12, dtype=np.uint8)}


@pytest.mark.parametrize("wpc,alpha", [(1, "dna"), (2, "dna"), (1, "wide"), (2, "wide"), (2, "nul8"), (2, "nul12"), (16, "dna")])
def test_random_shapes_all_phases(swb, oracle, wpc, alpha):
    # every pitch phase (m+1 mod 4), partial last strips, single rows/columns, sizes
    # straddling the 64-row strip, the 8-step group and the fast/edge switch
    # (warps_per_band > 2 is clamped to 2: the (16, "dna") case checks the clamp, not a new configuration)
    ACGT = ALPHABETS[alpha]
    rng = np.random.default_rng(100 + wpc)
    shapes = [(1, 1), (1, 77), (77, 1), (2, 3), (31, 32), (32, 33), (33, 31), (63, 65), (127, 129), (128, 128),
              (255, 31), (256, 256), (257, 259), (258, 64), (259, 97), (511, 300), (700, 513), (1025, 130),
              (1026, 257), (1027, 40), (1500, 700)]
    for (m, n) in shapes:
        a, b = rng.choice(ACGT, m), rng.choice(ACGT, n)
        H, P, maxPos, maxScore = gpu_fill(swb, a, b, wpc=wpc)
        Ho, Po, mpo = oracle.fill(a, b)
        assert (H == Ho).all(), (m, n, wpc, np.argwhere(H != Ho)[:4])
        assert (P == Po).all(), (m, n, wpc, np.argwhere(P != Po)[:4])
        assert maxPos == mpo and maxScore == Ho.max(), (m, n, wpc)


def test_tie_heavy_maxpos(swb, oracle):
    # 256x256 random DNA: a multi-cell global maximum in ~30% of the runs
    ties = 0
    for seed in range(60):
        a, b = oracle.generate(1000 + seed, 256, 256)
        H, P, maxPos, maxScore = gpu_fill(swb, a, b)
        Ho, Po, mpo = oracle.fill(a, b, order="wavefront")
        ties += int((Ho == Ho.max()).sum() > 1)
        assert maxPos == mpo and (H == Ho).all() and (P == Po).all(), seed
    assert ties >= 5


def test_scoring_alphabet_and_degenerate(swb, oracle):
    rng = np.random.default_rng(7)
    # no positive score anywhere -> maxPos 0, backtrack is a no-op
    a, b = np.full(300, ord("A"), np.uint8), np.full(200, ord("C"), np.uint8)
    H, P, maxPos, maxScore, plen = gpu_fill(swb, a, b, backtrack=True)
    assert H.max() == 0 and P.max() == 0 and P.min() == 0 and maxPos == 0 and maxScore == 0 and plen == 0
    # identical sequences: one long diagonal
    a = rng.choice(ACGT, 777)
    H, P, maxPos, maxScore, plen = gpu_fill(swb, a, a, backtrack=True)
    assert maxScore == 3 * 777 and maxPos == 778 * 777 + 777 and plen == 777
    # upstream's original scores (omp_smithW_orig.c:65-67) and a protein-like byte alphabet
    for scoring in [(5, -3, -4), (1, -1, -1), (2, -7, -1), (10, -2, -9)]:
        a = rng.integers(0, 256, 413, dtype=np.uint8); b = rng.integers(0, 256, 298, dtype=np.uint8)
        b[50:250] = a[100:300]                       # plant a long match
        H, P, maxPos, maxScore = gpu_fill(swb, a, b, scoring=scoring)
        Ho, Po, mpo = oracle.fill(a, b, scoring=scoring)
        assert (H == Ho).all() and (P == Po).all() and maxPos == mpo, scoring


def test_padded_pitch(swb, oracle):
    rng = np.random.default_rng(11)
    for (m, n, pitch) in [(100, 70, 104), (301, 99, 303), (258, 130, 320)]:
        a, b = rng.choice(ACGT, m), rng.choice(ACGT, n)
        H, P, maxPos, maxScore = gpu_fill(swb, a, b, pitch=pitch)
        Ho, Po, mpo = oracle.fill(a, b)
        assert (H == Ho).all() and (P == Po).all()
        i, j = divmod(mpo, m + 1)
        assert maxPos == i * pitch + j


def test_operator_and_host_api(swb, oracle):
    # smithWaterman(a,b,w,h,H,P,&maxloc) mirror + the host-buffer call
    a, b = swb.generate(42, 1000, 600)
    an, bn = np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8)
    Ho, Po, mpo = oracle.fill(an, bn)
    dH = torch.empty(601 * 1001, dtype=torch.int32, device="cuda:0")
    dP = torch.empty_like(dH)
    assert swb.smithWaterman(a, b, 1000, 600, dH, dP) == mpo
    assert (dH.view(601, 1001).cpu().numpy() == Ho).all()
    H = np.empty((601, 1001), np.int32); P = np.empty((601, 1001), np.int32)
    maxPos, plen = swb.align_host(a, b, H, P)
    leno = oracle.backtrack(Po, mpo)
    assert maxPos == mpo and plen == leno and (H == Ho).all() and (P == Po).all()
    with swb.AlignContext(1000, 600) as ctx:
        for _ in range(2):
            H[:] = -1; P[:] = -1
            assert ctx.align(a, b, H, P) == (mpo, leno)
            assert (H == Ho).all() and (P == Po).all()


def test_error_behaviour(swb):
    dH = torch.empty(64, dtype=torch.int32, device="cuda:0")
    with pytest.raises(swb.SwbError):
        swb.fill(b"ACGT", 4, b"ACG", 3, dH, dH, pitch=4)          # pitch < m+1
    with pytest.raises(swb.SwbError):
        swb.fill(b"ACGT", 4, b"ACG", 3, dH.data_ptr() + 4, dH, pitch=5)   # misaligned
    with pytest.raises(swb.SwbError):
        swb.fill(b"ACGT", 0, b"ACG", 3, dH, dH)
    # backtrack stages band rows with 16-byte bulk copies: a misaligned P is an error code, not a device fault
    with pytest.raises(swb.SwbError):
        swb.backtrack(dH.data_ptr() + 4, 5, 7)
    with pytest.raises(swb.SwbError):
        swb.backtrack_async(dH.data_ptr() + 8, 5, 7)
    torch.cuda.synchronize()


@pytest.mark.parametrize("cols,rows,seed", [(8192, 8192, 42), (20000, 3000, 5), (3000, 20000, 6)])
def test_medium_blockwise(swb, oracle, cols, rows, seed):
    a, b = oracle.generate(seed, cols, rows)
    dev = torch.device("cuda:0")
    dH = torch.empty((rows + 1) * (cols + 1), dtype=torch.int32, device=dev)
    dP = torch.empty_like(dH)
    maxPos = swb.fill(a, cols, b, rows, dH, dP)
    Hv, Pv = dH.view(rows + 1, cols + 1), dP.view(rows + 1, cols + 1)
    assert int(Hv[0].abs().sum()) == 0 and int(Pv[0].abs().sum()) == 0
    for i0, i1, Hb, Pb in oracle.fill_blocks(a, b, 1024):
        if i0 is None:
            assert maxPos == Pb
            break
        assert (Hv[i0:i1].cpu().numpy() == Hb).all(), (i0, i1)
        assert (Pv[i0:i1].cpu().numpy() == Pb).all(), (i0, i1)
    # backtrack vs oracle on the (host copy of) P
    Pfull = Pv.cpu().numpy().copy()
    leno = oracle.backtrack(Pfull, maxPos)
    assert swb.backtrack(dP, cols + 1, maxPos) == leno
    assert (Pv.cpu().numpy() == Pfull).all()


def test_score_only_kernel(swb, oracle):
    # no H/P stores: max score and maxPos (reference tie-break) from the per-row best cells
    rng = np.random.default_rng(21)
    for (m, n) in [(8, 9), (1, 1), (77, 130), (300, 64), (1027, 700), (4100, 1500), (700, 4100)]:
        a, b = rng.choice(ACGT, m), rng.choice(ACGT, n)
        ms, mp = swb.score_only(bytes(a), bytes(b))
        mso, mpo = oracle.score_only(a, b)
        assert (ms, mp) == (mso, mpo), (m, n)
    for seed in range(40):                      # tie-heavy shape
        a, b = oracle.generate(2000 + seed, 256, 256)
        assert swb.score_only(bytes(a), bytes(b)) == oracle.score_only(a, b), seed
    # no positive score
    assert swb.score_only(b"A" * 100, b"C" * 90) == (0, 0)


@pytest.mark.parametrize("m,n,npairs", [(256, 256, 96), (100, 70, 33), (513, 130, 17)])
def test_batch_of_pairs(swb, oracle, m, n, npairs):
    # BASELINE config 5 in small: independent equally shaped pairs in one launch
    rng = np.random.default_rng(m + n)
    A = rng.choice(ACGT, (npairs, m)); B = rng.choice(ACGT, (npairs, n))
    pitch = m + 1
    stride = ((n + 1) * pitch + 3) // 4 * 4
    dev = torch.device("cuda:0")
    dH = torch.full((npairs * stride,), -999, dtype=torch.int32, device=dev)
    dP = torch.full((npairs * stride,), -999, dtype=torch.int32, device=dev)
    d_pos = torch.zeros(npairs, dtype=torch.int64, device=dev)
    d_sc = torch.zeros(npairs, dtype=torch.int32, device=dev)
    swb.fill_batch_async(np.ascontiguousarray(A), m, np.ascontiguousarray(B), n, npairs, dH, dP, pitch, stride, d_pos, d_sc,
                         stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    Hh = dH.view(npairs, stride).cpu().numpy(); Ph = dP.view(npairs, stride).cpu().numpy()
    pos = d_pos.cpu().numpy(); scs = d_sc.cpu().numpy()
    for k in range(npairs):
        Ho, Po, mpo = oracle.fill(A[k], B[k])
        assert (Hh[k, :(n + 1) * pitch].reshape(n + 1, pitch) == Ho).all(), k
        assert (Ph[k, :(n + 1) * pitch].reshape(n + 1, pitch) == Po).all(), k
        assert pos[k] == mpo and scs[k] == Ho.max(), k
        assert (Hh[k, (n + 1) * pitch:] == -999).all()
    # score-only batch agrees
    d_pos2 = torch.zeros(npairs, dtype=torch.int64, device=dev); d_sc2 = torch.zeros(npairs, dtype=torch.int32, device=dev)
    swb.score_only_async(np.ascontiguousarray(A), m, np.ascontiguousarray(B), n, npairs, d_pos2, d_sc2,
                         stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert (d_pos2.cpu().numpy() == pos).all() and (d_sc2.cpu().numpy() == scs).all()
