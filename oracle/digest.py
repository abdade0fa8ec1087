"""Position-weighted 64-bit digests of sub-blocks of H and P (TEST INFRASTRUCTURE ONLY).

The multi-GB configurations (100000 x 100000 = 80 GB of H+P, 1000 x 2 000 000, ...) cannot be
compared element by element against a CPU run on the GPU box in reasonable time, and FNV-1a is
a serial byte hash.  These digests are sums, so they can be evaluated in parallel (torch on the
GPU, numpy on the CPU) and over any partition of the matrix into column strips:

    D1(X, rows i0..i1-1, cols j0..j1-1) = sum x[i][j] * w(i, j)            mod 2^64
    D2(...)                            = sum (x[i][j]^2 + 1) * w(j, i)    mod 2^64
    w(i, j) = 2 * (i * A + j * B) + 1                                      mod 2^64

(i, j) are GLOBAL row / column indices of the (n+1) x (m+1) matrix of the whole pair, so a GPU
that owns a column strip hashes its own part with its global column offset and the digests of
the strips of one chunk are comparable with the oracle's whatever the number of GPUs.
`oracle/make_golden_large.py` writes the oracle's digests to tests/golden/large_digests.json.
"""
from __future__ import annotations

import numpy as np

A = 0x9E3779B97F4A7C15
B = 0xC2B2AE3D27D4EB4F
_M = (1 << 64) - 1


def _s64(x: int) -> int:
    """two's-complement int64 view of a 64-bit pattern"""
    x &= _M
    return x - (1 << 64) if x >= (1 << 63) else x


def digest_np(X: np.ndarray, i0: int, j0: int) -> tuple[int, int]:
    """X: 2-D integer array holding rows i0.. and columns j0.. of the matrix."""
    with np.errstate(over="ignore"):
        x = X.astype(np.int64).view(np.uint64)
        ii = np.arange(i0, i0 + X.shape[0], dtype=np.uint64)[:, None]
        jj = np.arange(j0, j0 + X.shape[1], dtype=np.uint64)[None, :]
        a, b = np.uint64(A), np.uint64(B)
        w1 = (ii * a + jj * b) * np.uint64(2) + np.uint64(1)
        w2 = (jj * a + ii * b) * np.uint64(2) + np.uint64(1)
        d1 = int((x * w1).sum(dtype=np.uint64))
        d2 = int(((x * x + np.uint64(1)) * w2).sum(dtype=np.uint64))
    return d1, d2


def digest_torch(X, i0: int, j0: int) -> tuple[int, int]:
    """Same digests for a 2-D torch integer tensor (any device); int64 arithmetic wraps."""
    import torch
    x = X.to(torch.int64)
    ii = torch.arange(i0, i0 + X.shape[0], dtype=torch.int64, device=X.device)[:, None]
    jj = torch.arange(j0, j0 + X.shape[1], dtype=torch.int64, device=X.device)[None, :]
    a, b = _s64(A), _s64(B)
    w1 = (ii * a + jj * b) * 2 + 1
    w2 = (jj * a + ii * b) * 2 + 1
    d1 = int((x * w1).sum().item()) & _M
    d2 = int(((x * x + 1) * w2).sum().item()) & _M
    return d1, d2


def path_digest(positions) -> int:
    """positions: the GLOBAL linear indices i*(m+1)+j of the path cells (P < 0 after the backtrack),
    in any order.  -> sum over the ascending list of pos_k * (2k+1) * A  mod 2^64."""
    p = np.sort(np.asarray(positions, dtype=np.uint64))
    with np.errstate(over="ignore"):
        k = np.arange(p.size, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
        return int((p * k * np.uint64(A)).sum(dtype=np.uint64))
