"""Lane-vectorised numpy model of fill_kernel's schedule (smith-waterman_b200/csrc/swb_kernels.cuh).

It transliterates the kernel's index arithmetic -- blocks of four columns per lane and
step, the staging ring and the writer's per-row 128-byte segmentation (E, F), the tagged
strip -> strip hand-off ring (slot / epoch arithmetic, back-pressure rule), the tagged
band boundary rows and the loader -- with the 32 lanes of a warp as a numpy axis.

Compute warps, writer warps and loader warps are ACTORS that advance one group / round
at a time under a scheduling policy ("eager": consumers and writers first, "lazy":
producers run as far ahead as the back-pressure rules allow, writers as late as
possible, "random").  Poisoned rings make any read of a slot that was overwritten too
early, or not yet written, show up as a wrong result; a state in which no actor can run
is reported as a deadlock.  It checks the MATH and the flow-control RULES of the schedule,
not the memory model; the GPU tests check the real kernel.
"""
import numpy as np

TIE_NONE, TIE_DIAG, TIE_UP, TIE_LEFT = 8, 7, 5, 2
KT = 64
KR = 2                 # adjacent rows per lane (SWB_ROWS_PER_LANE)
STRIP_ROWS = 32 * KR
WRITERS = KR
ROW_INTS = 4 * KT
RING = 64
GROUP = 8
DRAIN_ROUNDS = 2
STAGE_SLACK = KT // GROUP - 2
POISON = -(1 << 40)


class Deadlock(AssertionError):
    pass


def fill_model(a, b, scoring=(3, -3, -2), wpc=2, pitch=None, policy="lazy", seed=0):
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    m, n = len(a), len(b)
    pitch = pitch or m + 1
    match, mismatch, gap = scoring
    sm, sx = 16 * match + TIE_DIAG, 16 * mismatch + TIE_DIAG
    gu, gl = 16 * gap + TIE_UP, 16 * gap + TIE_LEFT
    jmax = m >> 2
    ngroups = (jmax + 1 + 31 + GROUP - 1) // GROUP
    gtail = (jmax - 8) >> 3
    bstride = ngroups * GROUP
    strips = (n + STRIP_ROWS - 1) // STRIP_ROWS
    nbands = (strips + wpc - 1) // wpc
    lane = np.arange(32)
    rng = np.random.default_rng(seed)

    # packed a: chars of block j (columns 4j..4j+3, column c reads a[c-1])
    def achars(j):                      # j: (32,) -> (32, 4)
        c = 4 * j[:, None] + np.arange(4)[None, :]
        idx = c - 1
        ok = (idx >= 0) & (idx < m)
        out = np.zeros((32, 4), dtype=np.int64)
        out[ok] = a[idx[ok]]
        return out

    total = (n + 1) * pitch
    H = np.full(total, -777, dtype=np.int64)
    P = np.full(total, -777, dtype=np.int64)
    H[: m + 1] = 0
    P[: m + 1] = 0
    strip_max = np.zeros(strips, dtype=np.int64)
    boundary = np.zeros((max(nbands - 1, 0), bstride, 4), dtype=np.int64)     # memset 0 = invalid tags

    class Compute:
        def __init__(self, s):
            self.s = s
            self.band, self.w = divmod(s, wpc)
            self.r0 = 1 + STRIP_ROWS * s
            self.bch = []
            for q in range(KR):
                rows = self.r0 + KR * lane + q
                self.bch.append(np.where(rows <= n, b[np.minimum(rows, n) - 1], 0))
            self.A = np.zeros((32, 4), dtype=np.int64)
            self.dgp = np.zeros(32, dtype=np.int64)
            self.hl = [np.zeros(32, dtype=np.int64) for _ in range(KR)]
            self.g = 0
            self.started = False
            self.stage = np.full((STRIP_ROWS, ROW_INTS), POISON, dtype=np.int64)
            self.ring = np.zeros((RING, 4), dtype=np.int64)                  # my INPUT ring
            self.staged = 0
            self.drained = [0] * WRITERS
            self.consumed = 0                                                  # blocks of my input ring I am done with
            self.has_in = self.r0 > 1
            nxt = self.r0 + STRIP_ROWS <= n
            self.out = 0 if not nxt else (1 if self.w + 1 < wpc else 2)

        def ring_valid(self, j):
            e = self.ring[(j + 32) & (RING - 1)]
            want = 1 + (((j + 32) >> 6) & 1)
            return (e[0] & 3) == want

        def runnable(self):
            g = self.g
            if g >= ngroups:
                return False
            if g > STAGE_SLACK and min(self.drained) < g - STAGE_SLACK:
                return False
            t0 = g * GROUP
            if self.out == 1 and t0 - 80 > 0 and comp[self.s + 1].consumed < t0 - 80:
                return False
            if self.has_in:
                need = [0] if g == 0 else []
                need += [t + 1 for t in range(t0, t0 + GROUP) if t + 1 <= jmax]
                if not all(self.ring_valid(j) for j in need):
                    return False
            return True

        def take_input(self, j):
            assert self.ring_valid(j)
            e = self.ring[(j + 32) & (RING - 1)].copy()
            e[0] &= ~15
            self.A[0] = e

        def run(self):
            g = self.g
            t0 = g * GROUP
            if self.has_in:
                if g == 0:
                    self.take_input(0)
                self.consumed = t0
            fast = 4 <= g <= gtail
            for i in range(GROUP):
                t = t0 + i
                j = t - lane
                ch = achars(j)
                u = self.A.copy()
                dg = self.dgp.copy()
                self.dgp = self.A[:, 3].copy()
                slot = (t + lane) & (KT - 1)
                for q in range(KR):
                    sc = np.where(ch == self.bch[q][:, None], sm, sx)
                    hl = self.hl[q].copy()
                    dgn = self.hl[q].copy()                      # diagonal of the next row's first cell
                    k = np.zeros((32, 4), dtype=np.int64)
                    h = np.zeros((32, 4), dtype=np.int64)
                    for e in range(4):
                        ke = np.maximum(np.maximum(hl + gl, u[:, e] + gu), np.maximum(dg + sc[:, e], TIE_NONE))
                        if not fast:
                            ke = np.where((j < 0) | ((j == 0) & (e == 0)), TIE_NONE, ke)
                        else:
                            assert (j >= 1).all()
                        k[:, e] = ke
                        h[:, e] = ke & ~15
                        hl = h[:, e]
                        dg = u[:, e]
                    self.hl[q] = hl
                    dg = dgn
                    u = h
                    # stage: strip row KR*lane + q, slot (t + lane) & (KT-1)
                    for e in range(4):
                        self.stage[KR * lane + q, 4 * slot + e] = k[:, e]
                h = u                                             # the lane's last row
                # hand-off of lane 31
                j31 = t - 31
                if self.out and (fast or j31 >= 0):
                    blk = h[31].copy()
                    tag = 1 + (((t + 1) >> 6) & 1)
                    blk[0] |= tag
                    if self.out == 1:
                        comp[self.s + 1].ring[(t + 1) & (RING - 1)] = blk
                    else:
                        boundary[self.band, j31] = blk
                # row above for the next step
                newA = np.zeros((32, 4), dtype=np.int64)
                newA[1:] = h[:-1]
                self.A = newA
                if self.has_in and (fast or t + 1 <= jmax):
                    assert t + 1 <= jmax
                    self.take_input(t + 1)
            self.g += 1
            self.staged = self.g

    class Writer:
        def __init__(self, s, sub):
            self.s, self.sub = s, sub
            self.c = comp[s]
            r0 = self.c.r0
            rho = 32 * sub + lane
            self.rho = rho
            cl = rho // KR
            rows = r0 + rho
            ph = (rows * pitch) & 31
            d = (cl + ((31 - ph) >> 2)) >> 3
            self.E = 32 * d + ph
            self.G0 = rows * pitch - self.E
            assert (self.G0 % 32 == 0).all()
            self.F = (8 * cl - self.E) & (ROW_INTS - 1)
            self.rowok = rows <= n
            self.r = 0
            self.mx = 0
            self.rounds = ngroups + DRAIN_ROUNDS

        def runnable(self):
            return self.r < self.rounds and self.c.staged >= min(self.r + 1, ngroups)

        def run(self):
            r = self.r
            v = 32 * r + lane
            interior = self.rowok.all() and 32 * r - self.E.max() >= 0 and 32 * r + 31 - self.E.min() <= m
            for l in range(32):
                idx = (v + self.F[l]) & (ROW_INTS - 1)
                c = v - self.E[l]
                ok = np.full(32, True) if interior else (self.rowok[l] & (c >= 0) & (c <= m))
                if interior:
                    assert self.rowok[l] and (c >= 0).all() and (c <= m).all()
                k = self.c.stage[self.rho[l], idx]
                gi = self.G0[l] + v
                kk = k[ok]
                assert (kk != POISON).all(), "writer read a slot that was never written"
                assert (H[gi[ok]] == -777).all(), "cell written twice"
                H[gi[ok]] = kk >> 4
                P[gi[ok]] = kk & 3
                if kk.size:
                    self.mx = max(self.mx, int(kk.max()))
            self.r += 1
            self.c.drained[self.sub] = self.r
            if self.r == self.rounds:
                strip_max[self.s] = max(strip_max[self.s], self.mx >> 4)

    class Loader:
        def __init__(self, band):
            self.band = band
            self.c = comp[band * wpc]
            self.base = 0
            self.nblocks = jmax + 1

        def avail(self):
            limit = min(self.nblocks, self.c.consumed + RING)
            src = boundary[self.band - 1]
            n_ok = 0
            for j in range(self.base, min(limit, self.base + 32)):
                if (src[j, 0] & 3) == 1 + (((j + 32) >> 6) & 1):
                    n_ok += 1
                else:
                    break
            return n_ok

        def runnable(self):
            return self.base < self.nblocks and self.avail() > 0

        def run(self):
            src = boundary[self.band - 1]
            for j in range(self.base, self.base + self.avail()):
                self.c.ring[(j + 32) & (RING - 1)] = src[j].copy()
                self.base += 1

    comp = [Compute(s) for s in range(strips)]
    writers = [Writer(s, sub) for s in range(strips) for sub in range(WRITERS)]
    loaders = [Loader(bd) for bd in range(1, nbands)]

    def pick():
        run_c = [c for c in comp if c.runnable()]
        run_w = [w for w in writers if w.runnable()]
        run_l = [l for l in loaders if l.runnable()]
        if policy == "eager":            # drain first, consumers before producers
            order = run_w + run_l[::-1] + run_c[::-1]
        elif policy == "lazy":           # producers as far ahead as allowed, writers last
            order = run_c + run_l + run_w
        else:
            order = run_c + run_w + run_l
            if order:
                return order[rng.integers(len(order))]
        return order[0] if order else None

    while True:
        act = pick()
        if act is None:
            break
        act.run()
    done = all(c.g == ngroups for c in comp) and all(w.r == w.rounds for w in writers)
    if not done:
        raise Deadlock(f"stuck: compute {[c.g for c in comp]} of {ngroups}, writers {[w.r for w in writers]}")
    Hm = H.reshape(n + 1, pitch)[:, : m + 1]
    Pm = P.reshape(n + 1, pitch)[:, : m + 1]
    if pitch > m + 1:
        assert (H.reshape(n + 1, pitch)[:, m + 1:] == -777).all(), "padding written"
    return Hm, Pm, strip_max
