"""ctypes front-end to the CPU checkers (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package
(smith-waterman_b200/) must never do so.

* `Oracle`    -> oracle/libsworacle.so, our plain-C restatement of
                 omp_smithW.c (see sw_oracle.h for the per-function citations).
* `Reference` -> oracle/_ref/libswref.so, the UNMODIFIED reference translation
                 unit run in-process with a pinned seed (ref_shim.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "libsworacle.so"
REF_SO = HERE / "_ref" / "libswref.so"
REF_BIN = HERE / "_ref" / "omp_smithW_ref"
REF_V1_BIN = HERE / "_ref" / "v1_skipbt_ref"


def build(ref: bool = True) -> None:
    """(Re)build the checkers with oracle/Makefile.  Building is not using."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", str(HERE)] + targets, check=True)


class Scoring(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap", C.c_int32)]


DEFAULT_SCORING = (3, -3, -2)  # omp_smithW.c:75-77

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _as_bytes(s) -> np.ndarray:
    if isinstance(s, (bytes, bytearray)):
        return np.frombuffer(bytes(s), dtype=np.uint8).copy()
    if isinstance(s, str):
        return np.frombuffer(s.encode("ascii"), dtype=np.uint8).copy()
    return np.ascontiguousarray(s, dtype=np.uint8)


class Oracle:
    def __init__(self, path: Path = ORACLE_SO):
        if not path.exists():
            build(ref=False)
        self.lib = L = C.CDLL(str(path))
        L.swo_generate.argtypes = [C.c_uint, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.swo_generate.restype = None
        for name in ("swo_fill_wavefront", "swo_fill_rowmajor"):
            f = getattr(L, name)
            f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(Scoring),
                          _i32p, _i32p, C.POINTER(C.c_int64)]
            f.restype = None
        L.swo_backtrack.argtypes = [_i32p, C.c_int64, C.c_int64]
        L.swo_backtrack.restype = C.c_int64
        L.swo_fnv1a64.argtypes = [_i32p, C.c_int64]
        L.swo_fnv1a64.restype = C.c_uint64
        L.swo_score_only.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(Scoring),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.swo_score_only.restype = None
        L.swo_fill_block.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(Scoring),
                                     _i32p, _i32p, _i32p, C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_int64)]
        L.swo_fill_block.restype = None
        L.swo_nelement.argtypes = [C.c_int64] * 3
        L.swo_nelement.restype = C.c_int64

    def generate(self, seed: int, m: int, n: int):
        a = np.empty(max(m, 1), dtype=np.uint8)
        b = np.empty(max(n, 1), dtype=np.uint8)
        self.lib.swo_generate(seed, m, n, a.ctypes.data, b.ctypes.data)
        return a[:m], b[:n]

    def fill(self, a, b, scoring=DEFAULT_SCORING, order: str = "rowmajor"):
        """-> H, P ((n+1, m+1) int32, pre-backtrack), maxPos"""
        a = _as_bytes(a); b = _as_bytes(b)
        m, n = len(a), len(b)
        H = np.empty((n + 1, m + 1), dtype=np.int32)
        P = np.empty((n + 1, m + 1), dtype=np.int32)
        mp = C.c_int64(0)
        sc = Scoring(*scoring)
        fn = self.lib.swo_fill_wavefront if order == "wavefront" else self.lib.swo_fill_rowmajor
        fn(a.ctypes.data, m, b.ctypes.data, n, C.byref(sc), H.reshape(-1), P.reshape(-1), C.byref(mp))
        return H, P, int(mp.value)

    def fill_blocks(self, a, b, block_rows: int = 1024, scoring=DEFAULT_SCORING):
        """Generator over row blocks: yields (i0, i1, Hblk, Pblk) for rows i0..i1-1 (row 0
        excluded) and finally returns through .maxPos on the generator's last item:
        the last yield is (None, None, maxScore, maxPos)."""
        a = _as_bytes(a); b = _as_bytes(b)
        m, n = len(a), len(b)
        sc = Scoring(*scoring)
        top = np.zeros(m + 1, dtype=np.int32)
        best, bi, bj = C.c_int32(0), C.c_int64(0), C.c_int64(0)
        for i0 in range(1, n + 1, block_rows):
            i1 = min(i0 + block_rows, n + 1)
            Hb = np.empty((i1 - i0, m + 1), dtype=np.int32)
            Pb = np.empty((i1 - i0, m + 1), dtype=np.int32)
            self.lib.swo_fill_block(a.ctypes.data, m, b.ctypes.data, i0, i1, C.byref(sc), top,
                                    Hb.reshape(-1), Pb.reshape(-1), C.byref(best), C.byref(bi), C.byref(bj))
            top = Hb[-1].copy()
            yield i0, i1, Hb, Pb
        yield None, None, int(best.value), (int(bi.value) * (m + 1) + int(bj.value)) if best.value > 0 else 0

    def backtrack(self, P: np.ndarray, maxPos: int) -> int:
        """negates the path in place; returns its length"""
        return int(self.lib.swo_backtrack(P.reshape(-1), P.shape[1], maxPos))

    def score_only(self, a, b, scoring=DEFAULT_SCORING):
        a = _as_bytes(a); b = _as_bytes(b)
        ms, mp = C.c_int32(0), C.c_int64(0)
        sc = Scoring(*scoring)
        self.lib.swo_score_only(a.ctypes.data, len(a), b.ctypes.data, len(b), C.byref(sc),
                                C.byref(ms), C.byref(mp))
        return int(ms.value), int(mp.value)

    def fnv(self, x: np.ndarray) -> int:
        x = np.ascontiguousarray(x, dtype=np.int32)
        return int(self.lib.swo_fnv1a64(x.reshape(-1), x.size))


class Reference:
    """The unmodified reference omp_smithW.c, in-process, seed pinned."""

    def __init__(self, path: Path = REF_SO):
        if not path.exists():
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = L = C.CDLL(str(path))
        L.swref_run.argtypes = [C.c_longlong, C.c_longlong, C.c_uint, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.swref_run.restype = C.c_int

    def run(self, cols: int, rows: int, seed: int = 0, threads: int = 1, dump: bool = True):
        """cols <= 0 -> the built-in case.  Returns dict(a, b, H, P_bt, maxPos, path_len, t_fill, t_bt)."""
        m, n = (cols, rows) if cols > 0 else (8, 9)
        times = (C.c_double * 2)()
        if dump:
            H = np.zeros((n + 1, m + 1), dtype=np.int32)
            P = np.zeros((n + 1, m + 1), dtype=np.int32)
            a = np.zeros(m, dtype=np.uint8)
            b = np.zeros(n, dtype=np.uint8)
            rc = self.lib.swref_run(cols, rows, seed, threads, H.ctypes.data, P.ctypes.data,
                                    a.ctypes.data, b.ctypes.data, times)
        else:
            H = P = a = b = None
            rc = self.lib.swref_run(cols, rows, seed, threads, None, None, None, None, times)
        if rc != 0:
            raise RuntimeError(f"reference main returned {rc}")
        out = dict(t_fill=times[0], t_bt=times[1])
        if dump:
            neg = np.flatnonzero(P.reshape(-1) < 0)
            out.update(a=a, b=b, H=H, P_bt=P,
                       maxPos=int(neg.max()) if neg.size else 0, path_len=int(neg.size))
        return out


def run_reference_cli(cols: int, rows: int, threads: int | None, seed: int | None = 42,
                      binary: Path = REF_BIN, timeout: float = 3600.0):
    """Runs the stand-alone reference binary and parses its own timing lines.
    -> (fill_seconds, backtrack_seconds, threads_used)"""
    env = dict(os.environ)
    if threads is not None:
        env["OMP_NUM_THREADS"] = str(threads)
    if seed is not None:
        env["SWREF_SEED"] = str(seed)
    p = subprocess.run([str(binary), str(cols), str(rows)], env=env, capture_output=True, text=True,
                       timeout=timeout)
    t_fill = t_bt = None
    used = None
    for line in p.stdout.splitlines():
        if "Elapsed time for scoring matrix computation:" in line:
            t_fill = float(line.rsplit(":", 1)[1])
        elif "Elapsed time for backtracking:" in line:
            t_bt = float(line.rsplit(":", 1)[1])
        if line.startswith("Using ") and " out of max " in line:
            used = int(line.split()[1])
    if t_fill is None:
        # v1 prints "Elapsed time for scoring matrix computation" the same way; anything else is an error
        raise RuntimeError(f"reference binary gave no timing line (rc={p.returncode}):\n{p.stdout}\n{p.stderr}")
    return t_fill, t_bt, used
