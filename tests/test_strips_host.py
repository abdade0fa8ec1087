"""Host-side logic of the column-strip mode (smith-waterman_b200/strips.py) on CPU: the partition, the
maxPos reduction with the reference's tie-break and the right-to-left backtrack chain, checked against
the oracle; the cross-rank versions run over gloo with world_size 2 (no GPU involved: each rank's
"strip" is a numpy slice of the oracle's matrices with the hand-off marker in its local column 0)."""
import importlib
import os
import socket

import numpy as np
import pytest

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@pytest.fixture(scope="module")
def strips():
    importlib.import_module("smith-waterman_b200")
    return importlib.import_module("smith-waterman_b200.strips")


def local_view(H, P, c0, w, first):
    """what strip (c0, w) holds after its fill: local column 0 = copy of the left neighbour's last column,
    marked with the hand-off code in P"""
    Hl = H[:, c0:c0 + w + 1].copy()
    Pl = P[:, c0:c0 + w + 1].copy()
    if not first:
        Pl[1:, 0] = 5
        Pl[0, 0] = 0
    return Hl, Pl


def local_max(Hl, c0):
    """local maximum over local columns >= 1 with the reference's tie-break -> (score, i, j_global)"""
    sub = Hl[:, 1:]
    s = int(sub.max())
    if s <= 0:
        return (0, 0, 0)
    ii, jj = np.nonzero(sub == s)
    jj = jj + 1
    k = min(range(len(ii)), key=lambda t: (ii[t] + jj[t], -ii[t]))
    return (s, int(ii[k]), c0 + int(jj[k]))


def numpy_walk(Pl):
    """the backtrack kernel's contract on a local P (omp_smithW.c:405-420 + hand-off marker)"""
    def walk(i, j):
        n = 0
        while Pl[i, j] in (1, 2, 3):
            code = Pl[i, j]
            Pl[i, j] = -code
            n += 1
            if code == 3: i, j = i - 1, j - 1
            elif code == 1: i -= 1
            else: j -= 1
        return n, i, j
    return walk


def test_partition(strips):
    assert strips.partition(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert strips.partition(100000, 8) == [(12500 * g, 12500) for g in range(8)]
    for m, w in [(7, 7), (45001, 8), (13, 4)]:
        parts = strips.partition(m, w)
        assert parts[0][0] == 0 and sum(x[1] for x in parts) == m
        assert all(parts[g][0] + parts[g][1] == parts[g + 1][0] for g in range(w - 1))
        assert [strips.owner_of_column(parts, j) for j in (1, m)] == [0, w - 1]
    with pytest.raises(ValueError):
        strips.partition(3, 4)


@pytest.mark.parametrize("seed,m,n,world", [(1, 97, 80, 2), (2, 300, 120, 3), (3, 64, 257, 4), (4, 513, 77, 8), (5, 40, 40, 5)])
def test_maxpos_and_backtrack_chain_match_the_oracle(strips, oracle, seed, m, n, world):
    rng = np.random.default_rng(seed)
    a, b = rng.choice(ACGT, m), rng.choice(ACGT, n)
    b[n // 4: n // 4 + min(m, n) // 2] = a[m // 3: m // 3 + min(m, n) // 2]        # a path that crosses strips
    H, P, mp = oracle.fill(a, b, order="wavefront")
    parts = strips.partition(m, world)
    views = [local_view(H, P, c0, w, g == 0) for g, (c0, w) in enumerate(parts)]
    assert strips.reduce_maxpos([local_max(v[0], parts[g][0]) for g, v in enumerate(views)], m) == mp
    walkers = [numpy_walk(v[1]) for v in views]
    total, starts = strips.chain_backtrack(parts, mp, m, lambda g, i, j: walkers[g](i, j))
    Po = P.copy()
    assert total == oracle.backtrack(Po, mp)
    got = np.zeros_like(P)
    for g, (c0, w) in enumerate(parts):
        got[:, c0 + 1:c0 + w + 1] = views[g][1][:, 1:]
    assert (got == Po).all()
    assert len(starts) >= 1 and starts[0][0] == strips.owner_of_column(parts, mp % (m + 1))


def test_no_positive_score(strips):
    parts = strips.partition(20, 2)
    assert strips.reduce_maxpos([(0, 0, 0), (0, 0, 0)], 20) == 0
    assert strips.chain_backtrack(parts, 0, 20, lambda *x: 1 / 0) == (0, [])


def test_tie_break_across_strips(strips):
    # equal scores: the earlier anti-diagonal wins, then the larger row (omp_smithW.c:203-215,384-387)
    m = 100
    assert strips.reduce_maxpos([(9, 50, 10), (9, 20, 60)], m) == 50 * 101 + 10       # i+j = 60 < 80
    assert strips.reduce_maxpos([(9, 10, 50), (9, 5, 55)], m) == 10 * 101 + 50        # same diagonal: larger i
    assert strips.reduce_maxpos([(8, 1, 1), (9, 90, 99)], m) == 90 * 101 + 99


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, m, n, seed, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        from pathlib import Path
        sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
        importlib.import_module("smith-waterman_b200")
        st = importlib.import_module("smith-waterman_b200.strips")
        from oracle.swo import Oracle
        o = Oracle()
        rng = np.random.default_rng(seed)
        a, b = rng.choice(ACGT, m), rng.choice(ACGT, n)
        b[5:5 + n // 2] = a[m // 2 - n // 4: m // 2 - n // 4 + n // 2]
        H, P, mp = o.fill(a, b, order="wavefront")
        parts = st.partition(m, world)
        c0, w = parts[rank]
        Hl, Pl = local_view(H, P, c0, w, rank == 0)
        got_mp = st.allgather_maxpos(local_max(Hl, c0), m, dist)
        total = st.distributed_backtrack(parts, got_mp, m, rank, numpy_walk(Pl), dist)
        Po = P.copy(); want = o.backtrack(Po, mp)
        ok = got_mp == mp and total == want and (Pl[:, 1:] == Po[:, c0 + 1:c0 + w + 1]).all()
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_maxpos_and_backtrack(strips):
    mp = pytest.importorskip("torch.multiprocessing")
    ctx = mp.get_context("spawn")
    port = _free_port()
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 150, 90, 11, out)) for r in range(2)]
        for p in procs: p.start()
        for p in procs: p.join(120)
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: True, 1: True}
