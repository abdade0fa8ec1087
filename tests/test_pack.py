"""Packed transfer of H and P (include/swb200.h, swb_pack.cu): one byte per cell on the wire -- the row step of H
(gap .. match - gap by the recurrence, omp_smithW.c:331-388) and P (0..3, negated on the path, omp_smithW.c:405-420) --
expanded into the reference's int32 matrices (omp_smithW.c:203-216) on the host.  The host expander is tested here on
CPU against the format restated in numpy; the device packer and the whole swb_ctx_align path on the GPU."""
import numpy as np
import pytest


def np_pack(H, P, ppitch):
    """the wire format, restated: bits 7..3 = H[i][j] - H[i][j-1] + 16 (column -1 = 0), bits 2..0 = P + 3"""
    d = np.diff(H.astype(np.int64), axis=1, prepend=0) + 16
    assert d.min() >= 0 and d.max() <= 31 and P.min() >= -3 and P.max() <= 3
    out = np.zeros((H.shape[0], ppitch), np.uint8)
    out[:, : H.shape[1]] = (d.astype(np.uint8) << 3) | (P + 3).astype(np.uint8)
    return out


@pytest.mark.parametrize("rows,cols,threads", [(1, 1, 1), (3, 7, 2), (5, 8, 1), (17, 1001, 4), (64, 4097, 0), (2, 33, 8)])
def test_expand_rows_matches_the_format(swb, rows, cols, threads):
    rng = np.random.default_rng(rows * 1000 + cols)
    H = np.cumsum(rng.integers(-16, 16, (rows, cols)), axis=1).astype(np.int32)
    P = rng.integers(-3, 4, (rows, cols)).astype(np.int32)
    ppitch = swb.packed_pitch(cols)
    assert ppitch >= cols and ppitch % 64 == 0
    packed = np_pack(H, P, ppitch)
    for pitch in (cols, cols + 5):
        H2 = np.full((rows, pitch), -777, np.int32); P2 = np.full((rows, pitch), -777, np.int32)
        swb.expand_rows(packed, ppitch, rows, cols, H2, P2, pitch, None, threads)
        assert (H2[:, :cols] == H).all() and (P2[:, :cols] == P).all()
        assert (H2[:, cols:] == -777).all() and (P2[:, cols:] == -777).all()      # the padding stays untouched
        # either output may be left out
        H3 = np.full((rows, pitch), -777, np.int32)
        swb.expand_rows(packed, ppitch, rows, cols, H3, None, pitch, None, threads)
        assert (H3[:, :cols] == H).all()
        P3 = np.full((rows, pitch), -777, np.int32)
        swb.expand_rows(packed, ppitch, rows, cols, None, P3, pitch, None, threads)
        assert (P3[:, :cols] == P).all()


def test_expand_rows_unaligned_outputs(swb):
    # the vector body needs 32-byte aligned stores: every start offset must give the same answer
    rng = np.random.default_rng(5)
    rows, cols = 4, 259
    H = np.cumsum(rng.integers(-2, 6, (rows, cols)), axis=1).astype(np.int32)
    P = rng.integers(-3, 4, (rows, cols)).astype(np.int32)
    ppitch = swb.packed_pitch(cols)
    packed = np_pack(H, P, ppitch)
    for off_h in range(0, 9):
        for off_p in (0, 3):
            bufH = np.full(rows * cols + 16, -1, np.int32); bufP = np.full(rows * cols + 16, -1, np.int32)
            H2 = bufH[off_h: off_h + rows * cols].reshape(rows, cols); P2 = bufP[off_p: off_p + rows * cols].reshape(rows, cols)
            swb.expand_rows(packed, ppitch, rows, cols, H2, P2, cols, None, 1)
            assert (H2 == H).all() and (P2 == P).all()
            assert (bufH[:off_h] == -1).all() and (bufH[off_h + rows * cols:] == -1).all()


def test_oracle_matrices_fit_the_format(swb, oracle):
    # the claim the format rests on: gap <= H[i][j] - H[i][j-1] <= match - gap on real fills, path included
    for seed, m, n in ((42, 300, 200), (7, 64, 500)):
        a, b = swb.generate(seed, m, n)
        H, P, mp = oracle.fill(np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8))
        oracle.backtrack(P, mp)
        d = np.diff(H.astype(np.int64), axis=1, prepend=0)
        assert d.min() >= -2 and d.max() <= 5 and P.min() >= -3 and P.max() <= 3
        ppitch = swb.packed_pitch(m + 1)
        H2 = np.empty_like(H); P2 = np.empty_like(P)
        swb.expand_rows(np_pack(H, P, ppitch), ppitch, n + 1, m + 1, H2, P2, m + 1, None, 2)
        assert (H2 == H).all() and (P2 == P).all()


def test_expand_rows_with_row_base(swb):
    # a sub-matrix whose first column is not 0 (a column strip): column 0's step is stored as 0, the base carries it
    rng = np.random.default_rng(9)
    rows, cols = 13, 301
    base = rng.integers(0, 70000, rows).astype(np.int32)
    steps = rng.integers(-2, 6, (rows, cols)); steps[:, 0] = 0
    H = (base[:, None] + np.cumsum(steps, axis=1)).astype(np.int32)
    P = rng.integers(-3, 4, (rows, cols)).astype(np.int32)
    ppitch = swb.packed_pitch(cols)
    packed = np.zeros((rows, ppitch), np.uint8)
    packed[:, :cols] = ((steps + 16).astype(np.uint8) << 3) | (P + 3).astype(np.uint8)
    H2 = np.empty_like(H); P2 = np.empty_like(P)
    swb.expand_rows(packed, ppitch, rows, cols, H2, P2, cols, base, 3)
    assert (H2 == H).all() and (P2 == P).all()


def test_expand_rows_argument_errors(swb):
    buf = np.zeros(64, np.uint8); H = np.zeros(8, np.int32)
    with pytest.raises(swb.SwbError):
        swb.expand_rows(buf, 4, 1, 8, H, None, 8, None, 1)       # packed pitch < cols
    with pytest.raises(swb.SwbError):
        swb.expand_rows(buf, 64, 1, 8, H, None, 4, None, 1)      # pitch < cols
    with pytest.raises(swb.SwbError):
        swb.expand_rows(None, 64, 1, 8, H, None, 8, None, 1)


@pytest.mark.gpu
def test_device_pack_round_trip(swb, oracle):
    torch = pytest.importorskip("torch")
    for seed, m, n in ((42, 1000, 600), (3, 37, 1500), (9, 2049, 65)):
        a, b = swb.generate(seed, m, n)
        Ho, Po, mpo = oracle.fill(np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8))
        oracle.backtrack(Po, mpo)
        pitch = m + 1
        dH = torch.empty((n + 1) * pitch, dtype=torch.int32, device="cuda:0"); dP = torch.empty_like(dH)
        assert swb.smithWaterman(a, b, m, n, dH, dP) == mpo
        swb.backtrack(dP, pitch, mpo)
        ppitch = swb.packed_pitch(pitch)
        d_packed = torch.full(((n + 1) * ppitch,), 0xEE, dtype=torch.uint8, device="cuda:0")
        d_flag = torch.zeros(1, dtype=torch.int32, device="cuda:0")
        # in two row ranges, as a chunked caller would
        cut = (n + 1) // 3
        swb.pack_rows_async(dH, dP, pitch, 0, cut, pitch, d_packed, ppitch, d_flag, stream=torch.cuda.current_stream())
        swb.pack_rows_async(dH, dP, pitch, cut, n + 1 - cut, pitch, d_packed[cut * ppitch:], ppitch, d_flag,
                            stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert int(d_flag.item()) == 0
        packed = d_packed.cpu().numpy().reshape(n + 1, ppitch)
        assert (packed[:, :pitch] == np_pack(Ho, Po, ppitch)[:, :pitch]).all()
        H = np.empty((n + 1, pitch), np.int32); P = np.empty_like(H)
        swb.expand_rows(packed, ppitch, n + 1, pitch, H, P, pitch, None, 3)
        assert (H == Ho).all() and (P == Po).all()


@pytest.mark.gpu
def test_d2h_packed_sub_matrix_with_row_base(swb, oracle):
    # a block of columns of a real fill (what a column strip delivers): first column far from 0, own host pitch
    torch = pytest.importorskip("torch")
    m, n, c0, w = 2000, 900, 777, 1001
    a, b = swb.generate(5, m, n)
    Ho, Po, mpo = oracle.fill(np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8))
    oracle.backtrack(Po, mpo)
    pitch = m + 1
    dH = torch.empty((n + 1) * pitch, dtype=torch.int32, device="cuda:0"); dP = torch.empty_like(dH)
    assert swb.smithWaterman(a, b, m, n, dH, dP) == mpo
    swb.backtrack(dP, pitch, mpo)
    nbytes = swb.d2h_packed_scratch_bytes(n + 1, w)
    d_scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
    h_scratch = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hp = w + 3
    H = np.full((n + 1, hp), -7, np.int32); P = np.full((n + 1, hp), -7, np.int32)
    swb.d2h_packed(dH[c0:], dP[c0:], pitch, n + 1, w, H, P, hp, d_scratch, h_scratch, threads=3,
                   stream=torch.cuda.current_stream())
    assert (H[:, :w] == Ho[:, c0:c0 + w]).all() and (P[:, :w] == Po[:, c0:c0 + w]).all()
    assert (H[:, w:] == -7).all() and (P[:, w:] == -7).all()
    # values that do not fit: the same call delivers through the plain copies
    dH[5 * pitch + c0 + 10] += 1000
    Ho2 = Ho.copy(); Ho2[5, c0 + 10] += 1000
    swb.d2h_packed(dH[c0:], dP[c0:], pitch, n + 1, w, H, P, hp, d_scratch, h_scratch, stream=torch.cuda.current_stream())
    assert (H[:, :w] == Ho2[:, c0:c0 + w]).all() and (P[:, :w] == Po[:, c0:c0 + w]).all()
    assert (H[:, w:] == -7).all()


@pytest.mark.gpu
def test_device_pack_flags_values_that_do_not_fit(swb):
    torch = pytest.importorskip("torch")
    rows, cols = 4, 100
    ppitch = swb.packed_pitch(cols)
    for bad in ("step", "p"):
        dH = torch.zeros(rows * cols, dtype=torch.int32, device="cuda:0"); dP = torch.zeros_like(dH)
        if bad == "step": dH[2 * cols + 50:3 * cols] = 16          # a row step of +16
        else: dP[3 * cols + 7] = 4
        d_packed = torch.zeros(rows * ppitch, dtype=torch.uint8, device="cuda:0")
        d_flag = torch.zeros(1, dtype=torch.int32, device="cuda:0")
        swb.pack_rows_async(dH, dP, cols, 0, rows, cols, d_packed, ppitch, d_flag, stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert int(d_flag.item()) == 1


@pytest.mark.gpu
def test_ctx_align_packed_copy_back(swb, oracle):
    # 3000 x 3000 = 36 MB per matrix: above the 32 MiB threshold, so swb_ctx_align takes the packed path
    m = n = 3000
    a, b = swb.generate(11, m, n)
    Ho, Po, mpo = oracle.fill(np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8))
    leno = oracle.backtrack(Po, mpo)
    H = np.empty((n + 1, m + 1), np.int32); P = np.empty_like(H)
    with swb.AlignContext(m, n) as ctx:
        for _ in range(2):
            H[:] = -1; P[:] = -1
            assert ctx.align(a, b, H, P) == (mpo, leno)
            assert (H == Ho).all() and (P == Po).all()
        # a scoring whose row steps exceed the format: the flag sends the call down the plain copies
        wide = (40, -40, -30)
        Hw, Pw, mpw = oracle.fill(np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8), scoring=wide)
        lenw = oracle.backtrack(Pw, mpw)
        assert ctx.align(a, b, H, P, scoring=wide) == (mpw, lenw)
        assert (H == Hw).all() and (P == Pw).all()
