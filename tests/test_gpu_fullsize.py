"""BASELINE.json configs[1] (45000 x 45000, 16.2 GB of H+P) on one B200:
 * size-independent property: the recurrence itself, checked on the GPU for EVERY cell
   with plain torch ops (independent of the oracle and of the kernel's schedule),
 * the CPU oracle block by block (bit-exact H and P, maxPos with the reference tie-break),
 * backtrack: path cells negated, path is connected, ends at a NONE cell.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def check_recurrence_on_gpu(Hv, Pv, a_t, b_t, scoring, slab=512):
    """H[i][j] = max(0, H[i-1][j-1]+s, H[i-1][j]+g, H[i][j-1]+g) and P by the strict
    DIAGONAL, UP, LEFT order (omp_smithW.c:339-381), for all i, j >= 1."""
    match, mismatch, gap = scoring
    n, m = Hv.shape[0] - 1, Hv.shape[1] - 1
    assert int(Hv[:, 0].abs().sum()) == 0 and int(Hv[0].abs().sum()) == 0
    assert int(Pv[:, 0].abs().sum()) == 0 and int(Pv[0].abs().sum()) == 0
    for i0 in range(1, n + 1, slab):
        i1 = min(i0 + slab, n + 1)
        cur, up = Hv[i0:i1], Hv[i0 - 1:i1 - 1]
        sub = torch.where(b_t[i0 - 1:i1 - 1, None] == a_t[None, :], match, mismatch).to(torch.int32)
        d = up[:, :-1] + sub
        u = up[:, 1:] + gap
        l = cur[:, :-1] + gap
        zero = torch.zeros_like(d)
        m1 = torch.maximum(zero, d)
        m2 = torch.maximum(m1, u)
        h = torch.maximum(m2, l)
        pred = torch.where(l > m2, 2, torch.where(u > m1, 1, torch.where(d > 0, 3, 0))).to(torch.int32)
        assert bool((cur[:, 1:] == h).all()), f"H recurrence broken in rows {i0}..{i1}"
        assert bool((Pv[i0:i1, 1:] == pred).all()), f"P broken in rows {i0}..{i1}"


@pytest.mark.parametrize("cols,rows,seed", [(45000, 45000, 42)])
def test_fullsize_fill_and_backtrack(swb, oracle, cols, rows, seed):
    a, b = swb.generate(seed, cols, rows)
    dev = torch.device("cuda:0")
    cells = (rows + 1) * (cols + 1)
    dH = torch.empty(cells, dtype=torch.int32, device=dev)
    dP = torch.empty(cells, dtype=torch.int32, device=dev)
    maxPos = swb.fill(a, cols, b, rows, dH, dP)
    Hv, Pv = dH.view(rows + 1, cols + 1), dP.view(rows + 1, cols + 1)
    a_t = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev)
    b_t = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
    check_recurrence_on_gpu(Hv, Pv, a_t, b_t, (3, -3, -2))

    # maxPos: reference tie-break, recomputed independently on the GPU
    gmax = int(dH.max())
    idx = torch.nonzero(dH == gmax).flatten()
    i, j = idx // (cols + 1), idx % (cols + 1)
    key = (i + j) * (rows + 2) + (rows + 1 - i)          # smallest i+j, then largest i
    assert maxPos == int(idx[torch.argmin(key)])

    # the oracle, block by block, on the first and last 4096 rows and every 8th block between
    nblk = 0
    for k, (i0, i1, Hb, Pb) in enumerate(oracle.fill_blocks(a, b, 1024)):
        if i0 is None:
            assert Hb == gmax and Pb == maxPos
            break
        if i0 < 4096 or i1 > rows - 4096 or k % 8 == 0:
            assert (Hv[i0:i1].cpu().numpy() == Hb).all(), (i0, i1)
            assert (Pv[i0:i1].cpu().numpy() == Pb).all(), (i0, i1)
            nblk += 1
    assert nblk >= 12

    # backtrack (omp_smithW.c:405-420)
    before = dP.clone()
    plen = swb.backtrack(dP, cols + 1, maxPos)
    neg = torch.nonzero(dP < 0).flatten()
    assert plen == neg.numel() and plen > 0
    assert bool((dP[neg] == -before[neg]).all())
    changed = int((dP != before).sum())
    assert changed == plen
    # connected path from maxPos up/left to a NONE cell
    pos = neg.flip(0)                       # descending linear index = path order
    assert int(pos[0]) == maxPos
    step = torch.tensor([0, cols + 1, 1, cols + 2], device=dev)[before[pos].long()]
    assert bool((pos[:-1] - step[:-1] == pos[1:]).all())
    assert int(before[int(pos[-1]) - int(step[-1])]) == 0
    del before
