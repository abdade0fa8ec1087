"""smith-waterman_b200 -- host-side mirror of the reference's hot-path interface.

Thin ctypes layer over the C ABI in include/swb200.h (libswb200.so, hand-written
sm_100a CUDA).  There is no CPU fallback: importing works without a GPU (so the
symbol table can be checked), every compute call needs a CUDA device and raises
`SwbError` otherwise.  PyTorch is used by callers only for device memory and
streams; nothing here imports it.

Reference interfaces mirrored (chunhualiao/Smith-Waterman):
  smithWaterman(a, b, w, h, H, P, &maxloc)      rotated-cuda/sw-rotated-cuda-unified.cu:198-215
  similarityScore / nDiag loop, backtrack       omp_smithW.c:203-216,331-388,405-420
  generate()                                    omp_smithW.c:489-519
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

__all__ = ["lib", "SwbError", "Scoring", "DEFAULT_SCORING", "generate", "fill", "fill_async", "fill_batch_async", "score_only_async", "backtrack", "backtrack_async",
           "smithWaterman", "align_host", "AlignContext", "score_only", "KernelTimer", "host_alloc", "host_free",
           "device_count", "LIB_PATH", "NONE", "UP", "LEFT", "DIAGONAL", "PATH",
           "fill_pairs_async", "shard_pairs", "MultiGpuPair", "traceback_async", "cigar_from_moves", "alignment_from_moves", "read_sequences", "read_sequence", "load_manifest"]

# omp_smithW.c:32-36
PATH, NONE, UP, LEFT, DIAGONAL = -1, 0, 1, 2, 3

LIB_PATH = Path(os.environ.get("SWB_LIB", Path(__file__).resolve().parent / "libswb200.so"))


class SwbError(RuntimeError):
    pass


class Scoring(C.Structure):
    """omp_smithW.c:75-77"""
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap", C.c_int32)]


class Tuning(C.Structure):
    _fields_ = [("warps_per_band", C.c_int32), ("reserved", C.c_int32 * 5), ("timer", C.c_void_p),
                ("trace", C.c_void_p)]


DEFAULT_SCORING = (3, -3, -2)


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise SwbError(f"{LIB_PATH} is missing: build it with `make -C smith-waterman_b200` "
                       "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    L = C.CDLL(str(LIB_PATH))
    i64, i32, vp = C.c_int64, C.c_int32, C.c_void_p
    L.swb_strerror.argtypes = [C.c_int]; L.swb_strerror.restype = C.c_char_p
    L.swb_last_cuda_error.argtypes = []; L.swb_last_cuda_error.restype = C.c_char_p
    L.swb_version.restype = C.c_int
    L.swb_device_count.restype = C.c_int
    L.swb_fill_async.argtypes = [vp, i64, vp, i64, C.POINTER(Scoring), vp, vp, i64, vp, vp, C.c_int, vp,
                                 C.POINTER(Tuning)]
    L.swb_fill.argtypes = [vp, i64, vp, i64, C.POINTER(Scoring), vp, vp, i64, C.POINTER(i64), C.c_int, vp]
    L.swb_backtrack_async.argtypes = [vp, i64, i64, vp, vp, C.c_int, vp]
    L.swb_backtrack.argtypes = [vp, i64, i64, C.POINTER(i64), C.c_int, vp]
    L.swb_align_host.argtypes = [vp, i64, vp, i64, C.POINTER(Scoring), vp, vp, C.POINTER(i64), C.POINTER(i64),
                                 C.c_int, C.c_int]
    L.swb_ctx_create.argtypes = [C.POINTER(vp), i64, i64, C.c_int]
    L.swb_ctx_align.argtypes = [vp, vp, vp, C.POINTER(Scoring), vp, vp, C.POINTER(i64), C.POINTER(i64), C.c_int]
    L.swb_ctx_dH.argtypes = [vp]; L.swb_ctx_dH.restype = vp
    L.swb_ctx_dP.argtypes = [vp]; L.swb_ctx_dP.restype = vp
    L.swb_ctx_destroy.argtypes = [vp]; L.swb_ctx_destroy.restype = None
    L.swb_score_only.argtypes = [vp, i64, vp, i64, C.POINTER(Scoring), C.POINTER(i32), C.POINTER(i64), C.c_int, vp]
    L.swb_score_only_async.argtypes = [vp, i64, vp, i64, i64, C.POINTER(Scoring), vp, vp, C.c_int, vp, C.POINTER(Tuning)]
    L.swb_fill_batch_async.argtypes = [vp, i64, vp, i64, i64, C.POINTER(Scoring), vp, vp, i64, i64, vp, vp, C.c_int, vp,
                                       C.POINTER(Tuning)]
    L.swb_fill_strip_async.argtypes = [vp, i64, vp, i64, C.POINTER(Scoring), vp, vp, i64, vp, vp, vp, vp, i32, vp, vp,
                                       C.c_int, vp, C.POINTER(Tuning)]
    L.swb_strip_flag_count.argtypes = [i64]; L.swb_strip_flag_count.restype = i64
    L.swb_backtrack_from_async.argtypes = [vp, i64, i64, vp, vp, C.c_int, vp]
    L.swb_ipc_alloc.argtypes = [C.c_size_t, C.c_int]; L.swb_ipc_alloc.restype = vp
    L.swb_ipc_free.argtypes = [vp, C.c_int]; L.swb_ipc_free.restype = None
    L.swb_ipc_get_handle.argtypes = [vp, vp]
    L.swb_ipc_open.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.swb_ipc_close.argtypes = [vp, C.c_int]
    L.swb_enable_peer.argtypes = [C.c_int, C.c_int]
    L.swb_generate.argtypes = [C.c_uint, i64, i64, vp, vp]; L.swb_generate.restype = None
    L.swb_timer_create.argtypes = [C.POINTER(vp), C.c_int]; L.swb_timer_create.restype = C.c_int
    L.swb_timer_elapsed_ms.argtypes = [vp, C.POINTER(C.c_float)]; L.swb_timer_elapsed_ms.restype = C.c_int
    L.swb_timer_destroy.argtypes = [vp]; L.swb_timer_destroy.restype = None
    L.swb_host_alloc.argtypes = [C.c_size_t]; L.swb_host_alloc.restype = vp
    L.swb_host_free.argtypes = [vp]; L.swb_host_free.restype = None
    pi64 = C.POINTER(i64)
    L.swb_fill_pairs_async.argtypes = [vp, pi64, pi64, vp, pi64, pi64, pi64, i64, C.POINTER(Scoring), vp, vp, vp, vp,
                                       C.c_int, vp]
    L.swb_shard_pairs.argtypes = [i64, C.c_int, C.c_int, pi64, pi64]
    L.swb_multi_create.argtypes = [C.POINTER(vp), i64, i64, C.POINTER(C.c_int), C.c_int]
    L.swb_multi_fill.argtypes = [vp, vp, vp, C.POINTER(Scoring), pi64, C.POINTER(i32)]
    L.swb_multi_backtrack.argtypes = [vp, i64, pi64]
    L.swb_multi_align.argtypes = [vp, vp, vp, C.POINTER(Scoring), pi64, C.POINTER(i32), pi64, C.c_int]
    L.swb_multi_strips.argtypes = [vp]
    L.swb_multi_strip.argtypes = [vp, C.c_int, C.POINTER(C.c_int), pi64, pi64, pi64, C.POINTER(vp), C.POINTER(vp)]
    L.swb_multi_gather_host.argtypes = [vp, vp, vp]
    L.swb_multi_destroy.argtypes = [vp]; L.swb_multi_destroy.restype = None
    L.swb_fill_multi.argtypes = [vp, i64, vp, i64, C.POINTER(Scoring), C.POINTER(C.c_int), C.c_int, vp, vp, pi64, pi64,
                                 C.c_int]
    L.swb_seq_count.argtypes = [C.c_char_p, pi64]
    L.swb_seq_read.argtypes = [C.c_char_p, i64, C.POINTER(vp), pi64, C.c_char_p, C.c_size_t]
    L.swb_seq_free.argtypes = [vp]; L.swb_seq_free.restype = None
    L.swb_manifest_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.swb_manifest_pairs.argtypes = [vp]; L.swb_manifest_pairs.restype = i64
    L.swb_manifest_pair.argtypes = [vp, i64, C.POINTER(vp), pi64, C.POINTER(vp), pi64]
    L.swb_manifest_free.argtypes = [vp]; L.swb_manifest_free.restype = None
    L.swb_traceback_async.argtypes = [vp, i64, i64, vp, vp, vp, vp, C.c_int, vp]
    L.swb_cigar_from_moves.argtypes = [vp, i64, C.c_char_p, C.c_size_t]; L.swb_cigar_from_moves.restype = i64
    L.swb_alignment_from_moves.argtypes = [vp, i64, vp, vp, i64, i64, C.c_char_p, C.c_char_p]
    L.swb_traceback_async.restype = C.c_int; L.swb_alignment_from_moves.restype = C.c_int
    L.swb_packed_pitch.argtypes = [i64]; L.swb_packed_pitch.restype = i64
    L.swb_host_threads.argtypes = []; L.swb_host_threads.restype = C.c_int
    L.swb_pack_rows_async.argtypes = [vp, vp, i64, i64, i64, i64, vp, i64, vp, vp, C.c_int, vp]; L.swb_pack_rows_async.restype = C.c_int
    L.swb_expand_rows.argtypes = [vp, i64, i64, i64, vp, vp, i64, vp, C.c_int]; L.swb_expand_rows.restype = C.c_int
    L.swb_d2h_packed_scratch_bytes.argtypes = [i64, i64]; L.swb_d2h_packed_scratch_bytes.restype = C.c_size_t
    L.swb_d2h_packed.argtypes = [vp, vp, i64, i64, i64, vp, vp, i64, vp, vp, C.c_int, C.c_int, vp]; L.swb_d2h_packed.restype = C.c_int
    for name in ("swb_fill_pairs_async", "swb_shard_pairs", "swb_multi_create", "swb_multi_fill", "swb_multi_backtrack",
                 "swb_multi_align", "swb_multi_strips", "swb_multi_strip", "swb_multi_gather_host", "swb_fill_multi",
                 "swb_seq_count", "swb_seq_read", "swb_manifest_load", "swb_manifest_pair"):
        getattr(L, name).restype = C.c_int
    for name in ("swb_fill_async", "swb_fill", "swb_backtrack_async", "swb_backtrack", "swb_align_host",
                 "swb_ctx_create", "swb_ctx_align", "swb_score_only", "swb_score_only_async", "swb_fill_batch_async",
                 "swb_fill_strip_async", "swb_backtrack_from_async", "swb_ipc_get_handle", "swb_ipc_open", "swb_ipc_close",
                 "swb_enable_peer"):
        getattr(L, name).restype = C.c_int
    return L


lib = _load()


def _check(rc: int) -> None:
    if rc != 0:
        msg = lib.swb_strerror(rc).decode()
        detail = lib.swb_last_cuda_error().decode()
        raise SwbError(f"swb200: {msg} (status {rc})" + (f": {detail}" if rc in (-3, -5) and detail else ""))


def _scoring(sc) -> Scoring:
    return sc if isinstance(sc, Scoring) else Scoring(*(sc or DEFAULT_SCORING))


def _ptr(x) -> int:
    """raw address of a torch tensor / numpy array / int / bytes-like"""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    if isinstance(x, bytes):            # points into x itself; the caller keeps x alive over the call
        return C.cast(C.c_char_p(x), C.c_void_p).value
    if isinstance(x, bytearray):
        return C.addressof((C.c_char * len(x)).from_buffer(x))
    raise TypeError(f"cannot take the address of {type(x)}")


def _stream_ptr(stream) -> int:
    if stream is None:
        return 0
    return getattr(stream, "cuda_stream", stream)


def device_count() -> int:
    return int(lib.swb_device_count())


def generate(seed: int, m: int, n: int):
    """The reference's generate() for a pinned seed -> (a, b) as bytes."""
    a = C.create_string_buffer(max(m, 1))
    b = C.create_string_buffer(max(n, 1))
    lib.swb_generate(seed, m, n, a, b)
    return a.raw[:m], b.raw[:n]


def fill_async(a, m: int, b, n: int, dH, dP, pitch: int | None = None, d_maxPos=None, d_maxScore=None,
               scoring=None, device: int = 0, stream=None, warps_per_band: int = 0, timer=None, trace=None) -> None:
    """Enqueue the H/P/maxPos fill; a, b host or device; dH, dP device (n+1)*pitch int32."""
    sc = _scoring(scoring)
    tun = Tuning(warps_per_band=warps_per_band, timer=timer._h if timer is not None else None,
                 trace=_ptr(trace) if trace is not None else None)
    _check(lib.swb_fill_async(_ptr(a), m, _ptr(b), n, C.byref(sc), _ptr(dH), _ptr(dP), pitch or m + 1,
                              _ptr(d_maxPos), _ptr(d_maxScore), device, _stream_ptr(stream), C.byref(tun)))


def fill_batch_async(a, m: int, b, n: int, npairs: int, dH, dP, pitch: int | None = None, pair_stride: int | None = None,
                     d_maxPos=None, d_maxScore=None, scoring=None, device: int = 0, stream=None, warps_per_band: int = 0,
                     timer=None) -> None:
    """Enqueue the fill of npairs equally shaped pairs in one launch.  a: npairs*m bytes, b: npairs*n bytes
    (host or device); pair k's H/P start at dH/dP + k*pair_stride int32."""
    sc = _scoring(scoring)
    pitch = pitch or m + 1
    pair_stride = pair_stride or (n + 1) * pitch
    tun = Tuning(warps_per_band=warps_per_band, timer=timer._h if timer is not None else None)
    _check(lib.swb_fill_batch_async(_ptr(a), m, _ptr(b), n, npairs, C.byref(sc), _ptr(dH), _ptr(dP), pitch, pair_stride,
                                    _ptr(d_maxPos), _ptr(d_maxScore), device, _stream_ptr(stream), C.byref(tun)))


def score_only_async(a, m: int, b, n: int, npairs: int = 1, d_maxPos=None, d_maxScore=None, scoring=None,
                     device: int = 0, stream=None, warps_per_band: int = 0, timer=None, trace=None) -> None:
    """Enqueue the score-only kernel (no H/P stores) for npairs equally shaped pairs."""
    sc = _scoring(scoring)
    tun = Tuning(warps_per_band=warps_per_band, timer=timer._h if timer is not None else None,
                 trace=_ptr(trace) if trace is not None else None)
    _check(lib.swb_score_only_async(_ptr(a), m, _ptr(b), n, npairs, C.byref(sc), _ptr(d_maxPos), _ptr(d_maxScore), device,
                                    _stream_ptr(stream), C.byref(tun)))


def fill(a, m: int, b, n: int, dH, dP, pitch: int | None = None, scoring=None, device: int = 0, stream=None) -> int:
    """Blocking fill -> maxPos."""
    sc = _scoring(scoring)
    pos = C.c_int64(0)
    _check(lib.swb_fill(_ptr(a), m, _ptr(b), n, C.byref(sc), _ptr(dH), _ptr(dP), pitch or m + 1, C.byref(pos),
                        device, _stream_ptr(stream)))
    return int(pos.value)


def backtrack(dP, pitch: int, maxPos: int, device: int = 0, stream=None) -> int:
    """Negates the path in the device P in place -> path length."""
    n = C.c_int64(0)
    _check(lib.swb_backtrack(_ptr(dP), pitch, maxPos, C.byref(n), device, _stream_ptr(stream)))
    return int(n.value)


def backtrack_async(dP, pitch: int, maxPos: int = 0, d_maxPos=None, d_pathLen=None, device: int = 0,
                    stream=None) -> None:
    """Enqueue the backtrack; d_maxPos (device int64, e.g. fill_async's output) takes precedence over maxPos."""
    _check(lib.swb_backtrack_async(_ptr(dP), pitch, maxPos, _ptr(d_maxPos), _ptr(d_pathLen), device,
                                   _stream_ptr(stream)))


def smithWaterman(a, b, w: int, h: int, H, P, device: int = 0, stream=None, scoring=None) -> int:
    """Operator of the rotated variants: smithWaterman(a, b, w, h, H, P, &maxloc)
    (rotated-cuda/sw-rotated-cuda-unified.cu:198-215): a of length w (columns), b of
    length h (rows), row-major (h+1)*(w+1) outputs that need not be initialised.
    H, P are DEVICE buffers; returns the offset of the maximum in H (the reference
    returns a pointer into H)."""
    return fill(a, w, b, h, H, P, w + 1, scoring=scoring, device=device, stream=stream)


def align_host(a: bytes, b: bytes, H=None, P=None, scoring=None, do_backtrack: bool = True, device: int = 0):
    """Host-buffer call: H2D, fill, backtrack, D2H.  H, P: host int32 buffers of
    (len(b)+1)*(len(a)+1) (numpy / pinned torch) or None.  -> (maxPos, path_len)"""
    sc = _scoring(scoring)
    pos, plen = C.c_int64(0), C.c_int64(0)
    _check(lib.swb_align_host(_ptr(a), len(a), _ptr(b), len(b), C.byref(sc), _ptr(H), _ptr(P), C.byref(pos),
                              C.byref(plen), int(do_backtrack), device))
    return int(pos.value), int(plen.value)


class AlignContext:
    """Keeps the device matrices of one shape alive across host-buffer calls."""

    def __init__(self, m: int, n: int, device: int = 0):
        self.m, self.n, self.device = m, n, device
        self._h = C.c_void_p(0)
        _check(lib.swb_ctx_create(C.byref(self._h), m, n, device))

    def align(self, a, b, H=None, P=None, scoring=None, do_backtrack: bool = True):
        sc = _scoring(scoring)
        pos, plen = C.c_int64(0), C.c_int64(0)
        _check(lib.swb_ctx_align(self._h, _ptr(a), _ptr(b), C.byref(sc), _ptr(H), _ptr(P), C.byref(pos),
                                 C.byref(plen), int(do_backtrack)))
        return int(pos.value), int(plen.value)

    @property
    def dH(self) -> int:
        return lib.swb_ctx_dH(self._h)

    @property
    def dP(self) -> int:
        return lib.swb_ctx_dP(self._h)

    def close(self) -> None:
        if self._h:
            lib.swb_ctx_destroy(self._h)
            self._h = C.c_void_p(0)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KernelTimer:
    """CUDA-event pair that swb_fill_async records around the fill kernel alone."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p(0)
        _check(lib.swb_timer_create(C.byref(self._h), device))

    def elapsed_ms(self) -> float:
        ms = C.c_float(0)
        _check(lib.swb_timer_elapsed_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def close(self) -> None:
        if self._h:
            lib.swb_timer_destroy(self._h)
            self._h = C.c_void_p(0)


def score_only(a: bytes, b: bytes, scoring=None, device: int = 0, stream=None):
    sc = _scoring(scoring)
    ms, pos = C.c_int32(0), C.c_int64(0)
    _check(lib.swb_score_only(_ptr(a), len(a), _ptr(b), len(b), C.byref(sc), C.byref(ms), C.byref(pos), device,
                              _stream_ptr(stream)))
    return int(ms.value), int(pos.value)


def host_alloc(nbytes: int) -> int:
    p = lib.swb_host_alloc(nbytes)
    if not p:
        raise SwbError(f"pinned host allocation of {nbytes} bytes failed")
    return p


def host_free(p: int) -> None:
    lib.swb_host_free(p)


# ---------------------------------------------------------------------------------------------
# variable-length batches, pair-wise sharding (SURVEY 8(b) swb_fill_batch with m[], n[]; 8(e))
# ---------------------------------------------------------------------------------------------
def _i64_array(xs):
    return (C.c_int64 * len(xs))(*[int(x) for x in xs])


def fill_pairs_async(a, a_off, m, b, b_off, n, hp_off, dH, dP, d_maxPos=None, d_maxScore=None, scoring=None,
                     device: int = 0, stream=None) -> None:
    """Enqueue the fill of len(m) independent pairs of any shapes (swb_fill_pairs_async): a, b are host or
    device concatenations, a_off / b_off / hp_off / m / n host sequences of ints."""
    sc = _scoring(scoring)
    _check(lib.swb_fill_pairs_async(_ptr(a), _i64_array(a_off), _i64_array(m), _ptr(b), _i64_array(b_off), _i64_array(n),
                                    _i64_array(hp_off), len(m), C.byref(sc), _ptr(dH), _ptr(dP), _ptr(d_maxPos),
                                    _ptr(d_maxScore), device, _stream_ptr(stream)))


def shard_pairs(npairs: int, nshards: int, shard: int):
    """-> (first, count): the contiguous block of pairs GPU `shard` of `nshards` owns (swb_shard_pairs)."""
    first, count = C.c_int64(0), C.c_int64(0)
    _check(lib.swb_shard_pairs(npairs, nshards, shard, C.byref(first), C.byref(count)))
    return int(first.value), int(count.value)


class MultiGpuPair:
    """ONE pair in column strips over several GPUs of this process (swb_multi_*: the C++ host driver)."""

    def __init__(self, m: int, n: int, devices):
        self.m, self.n, self.devices = m, n, list(devices)
        self._h = C.c_void_p(0)
        arr = (C.c_int * len(self.devices))(*self.devices)
        _check(lib.swb_multi_create(C.byref(self._h), m, n, arr, len(self.devices)))

    def fill(self, a: bytes, b: bytes, scoring=None):
        """-> (maxPos, maxScore); maxPos indexes the (n+1) x (m+1) matrix of the whole pair"""
        sc = _scoring(scoring)
        pos, score = C.c_int64(0), C.c_int32(0)
        _check(lib.swb_multi_fill(self._h, _ptr(a), _ptr(b), C.byref(sc), C.byref(pos), C.byref(score)))
        return int(pos.value), int(score.value)

    def backtrack(self, maxPos: int) -> int:
        plen = C.c_int64(0)
        _check(lib.swb_multi_backtrack(self._h, maxPos, C.byref(plen)))
        return int(plen.value)

    def strip(self, g: int):
        """-> dict(device, col0, m, pitch, dH, dP) of strip g (dH, dP raw device addresses)"""
        dev = C.c_int(0); col0, ml, pitch = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        dH, dP = C.c_void_p(0), C.c_void_p(0)
        _check(lib.swb_multi_strip(self._h, g, C.byref(dev), C.byref(col0), C.byref(ml), C.byref(pitch), C.byref(dH), C.byref(dP)))
        return dict(device=dev.value, col0=col0.value, m=ml.value, pitch=pitch.value, dH=dH.value, dP=dP.value)

    def gather_host(self, H=None, P=None) -> None:
        """copies the strips into host (n+1) x (m+1) int32 matrices (numpy / pinned); either may be None"""
        _check(lib.swb_multi_gather_host(self._h, _ptr(H), _ptr(P)))

    def close(self) -> None:
        if self._h:
            lib.swb_multi_destroy(self._h)
            self._h = C.c_void_p(0)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ---------------------------------------------------------------------------------------------
# real sequence input: FASTA / UCSC .2bit files and batch manifests
# ---------------------------------------------------------------------------------------------
def read_sequence(path, record: int = 0):
    """-> (name, sequence bytes) of one record of a FASTA / .2bit file (swb_seq_read)"""
    seq, ln = C.c_void_p(0), C.c_int64(0)
    name = C.create_string_buffer(256)
    _check(lib.swb_seq_read(str(path).encode(), record, C.byref(seq), C.byref(ln), name, 256))
    try:
        return name.value.decode(), C.string_at(seq.value, ln.value)
    finally:
        lib.swb_seq_free(seq)


def read_sequences(path):
    cnt = C.c_int64(0)
    _check(lib.swb_seq_count(str(path).encode(), C.byref(cnt)))
    return [read_sequence(path, k) for k in range(cnt.value)]


def load_manifest(path):
    """-> [(a bytes, b bytes)] of a batch manifest (swb_manifest_load)"""
    h = C.c_void_p(0)
    _check(lib.swb_manifest_load(str(path).encode(), C.byref(h)))
    try:
        out = []
        for k in range(lib.swb_manifest_pairs(h)):
            pa, pb, la, lb = C.c_void_p(0), C.c_void_p(0), C.c_int64(0), C.c_int64(0)
            _check(lib.swb_manifest_pair(h, k, C.byref(pa), C.byref(la), C.byref(pb), C.byref(lb)))
            out.append((C.string_at(pa.value, la.value), C.string_at(pb.value, lb.value)))
        return out
    finally:
        lib.swb_manifest_free(h)


# ---------------------------------------------------------------------------------------------
# alignment emission
# ---------------------------------------------------------------------------------------------
def traceback_async(dP, pitch: int, startPos: int = 0, d_startPos=None, d_pathLen=None, d_endPos=None, d_moves=None,
                    device: int = 0, stream=None) -> None:
    """backtrack that also writes the path's moves (device uint8 buffer, walk order; 1 UP, 2 LEFT, 3 DIAGONAL)"""
    _check(lib.swb_traceback_async(_ptr(dP), pitch, startPos, _ptr(d_startPos), _ptr(d_pathLen), _ptr(d_endPos),
                                   _ptr(d_moves), device, _stream_ptr(stream)))


def packed_pitch(cols: int) -> int:
    return int(lib.swb_packed_pitch(cols))


def host_threads() -> int:
    return int(lib.swb_host_threads())


def pack_rows_async(dH, dP, pitch: int, row0: int, nrows: int, cols: int, d_packed, packed_pitch_: int, d_overflow,
                    d_row_base=None, device: int = 0, stream=None) -> None:
    """DEVICE: one byte per cell (row step of H + P) of rows row0 .. row0+nrows-1 (include/swb200.h, packed transfer)."""
    _check(lib.swb_pack_rows_async(_ptr(dH), _ptr(dP), pitch, row0, nrows, cols, _ptr(d_packed), packed_pitch_,
                                   _ptr(d_overflow), _ptr(d_row_base), device, _stream_ptr(stream)))


def expand_rows(packed, packed_pitch_: int, nrows: int, cols: int, H, P, pitch: int, row_base=None, threads: int = 0) -> None:
    """HOST: packed rows -> int32 H and/or P (bit-exact)."""
    _check(lib.swb_expand_rows(_ptr(packed), packed_pitch_, nrows, cols, _ptr(H), _ptr(P), pitch, _ptr(row_base), threads))


def d2h_packed_scratch_bytes(nrows: int, cols: int) -> int:
    return int(lib.swb_d2h_packed_scratch_bytes(nrows, cols))


def d2h_packed(dH, dP, pitch: int, nrows: int, cols: int, H, P, host_pitch: int, d_scratch, h_scratch, threads: int = 0,
               device: int = 0, stream=None) -> None:
    """Device H/P -> HOST int32 H/P through the packed transfer (synchronous; plain copies if the format does not fit)."""
    _check(lib.swb_d2h_packed(_ptr(dH), _ptr(dP), pitch, nrows, cols, _ptr(H), _ptr(P), host_pitch, _ptr(d_scratch),
                              _ptr(h_scratch), threads, device, _stream_ptr(stream)))


def cigar_from_moves(moves: bytes) -> str:
    """moves in walk order -> CIGAR in sequence order (a = reference, b = query: M / I (UP) / D (LEFT))"""
    need = lib.swb_cigar_from_moves(_ptr(moves), len(moves), None, 0)
    if need < 0:
        _check(int(need))
    buf = C.create_string_buffer(need + 1)
    lib.swb_cigar_from_moves(_ptr(moves), len(moves), buf, need + 1)
    return buf.value.decode()


def alignment_from_moves(moves: bytes, a: bytes, b: bytes, startPos: int, pitch: int):
    """-> (gapped a, gapped b) as bytes"""
    oa, ob = C.create_string_buffer(len(moves) + 1), C.create_string_buffer(len(moves) + 1)
    _check(lib.swb_alignment_from_moves(_ptr(moves), len(moves), _ptr(a), _ptr(b), startPos, pitch, oa, ob))
    return oa.value, ob.value
