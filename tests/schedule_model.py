"""Lane-vectorised numpy model of fill_kernel's schedule (swb_kernels.cuh).

It transliterates the kernel's index arithmetic -- per-row 16-byte block phase,
lane skew sigma, the 8-register window of the row above, the staging ring and the
write-out addressing, the strip -> strip hand-off indices -- with the 32 lanes of a
warp as a numpy axis.  Strips run one after the other (no concurrency), so it checks
the MATH of the schedule, not the synchronisation.  tests/test_schedule_model.py
compares it with the oracle; the GPU tests compare the real kernel.
"""
import numpy as np

TIE_NONE, TIE_DIAG, TIE_UP, TIE_LEFT = 8, 7, 5, 2
RING = 64
AOFF = 64


def build_a4(a: np.ndarray, stride: int) -> np.ndarray:
    m = len(a)
    out = np.zeros((4, stride, 4), dtype=np.int64)
    for s in range(4):
        k = np.arange(stride)
        for e in range(4):
            idx = 4 * (k - AOFF) + s - 4 + e
            ok = (idx >= 0) & (idx < m)
            out[s, ok, e] = a[idx[ok]]
    return out            # bytes of each word, little-endian order e = 0..3


def fill_model(a, b, scoring=(3, -3, -2), wpc=4, pitch=None, check_fast=True):
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    m, n = len(a), len(b)
    pitch = pitch or m + 1
    MU = pitch & 3
    match, mismatch, gap = scoring
    sm, sx = 16 * match + TIE_DIAG, 16 * mismatch + TIE_DIAG
    gu, gl = 16 * gap + TIE_UP, 16 * gap + TIE_LEFT
    qbmax = ((m + 3) >> 2) + 1
    steps = (qbmax + 62 + 7) // 8 * 8
    stride = AOFF + steps + 8
    a4 = build_a4(a, stride)
    total = (n + 1) * pitch + 8
    H = np.full(total, -777, dtype=np.int64)
    P = np.full(total, -777, dtype=np.int64)
    H[: m + 1] = 0
    P[: m + 1] = 0
    row_max = np.zeros(n + 1, dtype=np.int64)
    lane = np.arange(32)
    strips = (n + 31) // 32
    ring_next = None          # blocks written by the previous strip's lane 31 (dict qb -> 4 ints)
    for s_idx in range(strips):
        w = s_idx % wpc
        r0 = 1 + 32 * s_idx
        phi0 = (r0 * pitch) & 3
        lm = phi0 + lane * MU
        sigma, phil = lm >> 2, lm & 3
        sigma31 = (phi0 + 31 * MU) >> 2
        wrap0 = 1 if phi0 < MU else 0
        cbase = -lane * (4 + MU) - phi0
        row = r0 + lane
        row_ok = row <= n
        bch = np.where(row_ok, b[np.minimum(row, n) - 1], 0)
        acopy = 3 - phil
        aoff = AOFF - lane - sigma
        has_consumer = (w + 1 < wpc) and (r0 + 32 <= n)
        src_global = w == 0
        # ---- input ring contents
        phi_prod = (phi0 - MU) & 3
        qbp = ((m + phi_prod) >> 2) + 1
        if src_global:
            base_blk = ((r0 - 1) * pitch) >> 2

            def ring_in(idx):
                if idx < qbp:
                    return H[4 * (base_blk + idx): 4 * (base_blk + idx) + 4] << 4
                return np.zeros(4, dtype=np.int64)
        else:
            prev = ring_next

            def ring_in(idx, prev=prev):
                return prev.get(idx, np.full(4, 123456, dtype=np.int64))   # stale garbage
        ring_out = {}
        A = np.zeros((32, 4), dtype=np.int64)
        B = np.zeros((32, 4), dtype=np.int64)
        A[0] = ring_in(wrap0)
        if wrap0:
            B[0] = ring_in(0)
        hl = np.zeros(32, dtype=np.int64)
        rmax = np.zeros(32, dtype=np.int64)
        stage = np.zeros((32, 8, 4), dtype=np.int64)
        t_lo = (1 + 31 * (4 + MU) + phi0 + 3) >> 2
        t_hi = (m - 3 + phi0) >> 2
        fk, fe = lane >> 3, lane & 7
        Z = r0 * pitch - phi0
        assert Z % 4 == 0
        q4 = (pitch - 4 - MU) >> 2
        g0 = (Z >> 2) - 7 + fe + (8 * fk) * q4
        fcol0 = 4 * (fe - 7) - (8 * fk) * (4 + MU) - phi0
        for t in range(steps):
            tg = t - (t % 8)
            fast = (tg - 7 >= t_lo) and (tg + 7 <= t_hi)
            aw = a4[acopy, aoff + t]                        # (32, 4) bytes
            mis = aw != bch[:, None]
            sc = np.where(mis, sx, sm)
            W = np.concatenate([B, A], axis=1)
            dg = W[:, 3 - MU]
            up = W[:, 4 - MU: 8 - MU]
            K = np.zeros((32, 4), dtype=np.int64)
            h = np.zeros((32, 4), dtype=np.int64)
            left = hl.copy()
            diag = dg
            for e in range(4):
                k = np.maximum(left + gl, np.maximum(up[:, e] + gu, np.maximum(diag + sc[:, e], TIE_NONE)))
                c = 4 * t + cbase + e
                valid = (c >= 1) & (c <= m)
                if fast and check_fast:
                    assert valid.all(), (s_idx, t, e)
                k = np.where(valid, k, TIE_NONE)
                K[:, e] = k
                h[:, e] = k & ~15
                left = h[:, e]
                diag = up[:, e]
            hl = h[:, 3]
            rmax = np.maximum(rmax, K.max(axis=1))
            stage[lane, (t + lane) & 7] = K
            qb31 = t - 31 - sigma31
            if has_consumer and qb31 >= 0:
                ring_out[qb31] = h[31].copy()
            B = A.copy()
            A = np.roll(h, 1, axis=0)
            A[0] = ring_in(t + 1 + wrap0)
            # ---- write-out
            c = (t + 1) & 7
            lk = c + 8 * fk
            kv = stage[lk, (2 * (t + 1) + fe) & 7]
            g = g0 + t + c * q4
            col = fcol0 + 4 * t - c * (4 + MU)
            rowmask = (r0 + lk) <= n
            for e in range(4):
                ok = rowmask & (col + e >= 0) & (col + e <= m)
                if fast and check_fast:
                    assert (ok == rowmask).all()
                idx = 4 * g[ok] + e
                # every element is written exactly once
                assert (H[idx] == -777).all(), (s_idx, t, e)
                H[idx] = kv[ok, e] >> 4
                P[idx] = kv[ok, e] & 3
        row_max[row[row_ok]] = (rmax >> 4)[row_ok]
        ring_next = ring_out
    Hm = H[: (n + 1) * pitch].reshape(n + 1, pitch)[:, : m + 1]
    Pm = P[: (n + 1) * pitch].reshape(n + 1, pitch)[:, : m + 1]
    return Hm, Pm, row_max
