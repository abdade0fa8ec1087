"""Column-strip mode: ONE Smith-Waterman pair across several GPUs (BASELINE config
"100000x100000 single pair, column-strip wavefront pipelined across 2/4/8 B200").

The reference has no multi-GPU form; this extends its nDiag wavefront (omp_smithW.c:203-216) the way
the north star asks: GPU g owns a contiguous block of columns of every row; the fill kernel of GPU g
pushes the H values of its last column into GPU g+1's memory with peer stores while it runs and
GPU g+1's kernel waits on per-piece flags (swb_fill_strip_async in include/swb200.h).  There is no
collective in the data path.  What crosses ranks on the host side is small:

  * maxPos    one all-gather of (score, i, j) per rank, reduced with the reference's tie-break
              (omp_smithW.c:203-215,384-387: first cell in anti-diagonal order, bottom-left first);
  * backtrack (omp_smithW.c:405-420) walks right to left: a strip's walk ends on NONE or on the
              hand-off marker of its local column 0, and continues in the last column of the strip
              on its left (at most world-1 hand-offs, one broadcast each).

Two drivers share the same strip object:
  StripSet       one process, strips on one or several devices (tests; peer access inside the process)
  StripPipeline  one process per GPU over torch.distributed (NCCL or gloo for the small host-side
                 exchanges), boundary buffers mapped through CUDA IPC

torch is imported lazily: the package itself stays importable without it.
"""
from __future__ import annotations

import ctypes as C

from . import lib, _check, _ptr, _scoring, _stream_ptr, Tuning, SwbError

HANDOFF = 5          # P code of local column 0 of a strip with a left neighbour (swb_kernels.cuh kHandOff)


# ----------------------------------------------------------------------------------------------
# pure host logic (no GPU, no torch): partition, maxPos reduction, backtrack chain
# ----------------------------------------------------------------------------------------------
def partition(m: int, world: int):
    """Contiguous column blocks: -> [(col0, m_local)] with global columns col0+1 .. col0+m_local."""
    if world < 1 or m < world:
        raise ValueError(f"cannot split {m} columns over {world} strips")
    base, rem = divmod(m, world)
    out, c = [], 0
    for g in range(world):
        w = base + (1 if g < rem else 0)
        out.append((c, w))
        c += w
    return out


def reduce_maxpos(entries, m: int) -> int:
    """entries: per strip (score, i, j_global) of its local maximum (score 0 = none).  The reference's
    running arg-max keeps the FIRST cell, in its scan order, that attains the global maximum: smallest
    i+j, then largest i (omp_smithW.c:203-215,282-291,384-387).  -> index in the (n+1) x (m+1) matrix, 0 if none."""
    best = None
    for (score, i, j) in entries:
        if score <= 0:
            continue
        key = (-score, i + j, -i)
        if best is None or key < best[0]:
            best = (key, i, j)
    return 0 if best is None else best[1] * (m + 1) + best[2]


def owner_of_column(parts, j: int) -> int:
    for g, (c0, w) in enumerate(parts):
        if c0 + 1 <= j <= c0 + w:
            return g
    raise ValueError(f"column {j} is not owned by any strip")


def chain_backtrack(parts, maxPos: int, m: int, walk):
    """Drives the right-to-left backtrack.  walk(g, i, j_local) -> (path_len, end_i, end_j_local) runs the walk
    of strip g from its local cell (i, j_local).  -> (total path length, [(g, i, j_local)] start cells used)."""
    if maxPos <= 0:
        return 0, []
    i, j = divmod(maxPos, m + 1)
    g = owner_of_column(parts, j)
    jl = j - parts[g][0]
    total, starts = 0, []
    while True:
        starts.append((g, i, jl))
        plen, ei, ej = walk(g, i, jl)
        total += plen
        if g > 0 and ej == 0 and plen >= 0 and ei >= 1:
            # ended on the hand-off marker: the same cell is the last column of the strip on the left
            g -= 1
            i, jl = ei, parts[g][1]
        else:
            return total, starts


# ----------------------------------------------------------------------------------------------
# one strip
# ----------------------------------------------------------------------------------------------
class ColumnStrip:
    """The share of one GPU: columns col0+1 .. col0+m_local of the pair (a, b)."""

    def __init__(self, a_local, b, col0: int, m_total: int, device: int, first: bool, scoring=None):
        import torch
        self.torch = torch
        self.device, self.first = device, first
        self.col0, self.m, self.m_total = col0, len(a_local), m_total
        self.n = len(b)
        self.pitch = self.m + 1
        self.scoring = _scoring(scoring)
        dev = torch.device("cuda", device)
        self.a_d = torch.frombuffer(bytearray(a_local), dtype=torch.uint8).to(dev)
        self.b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
        cells = (self.n + 1) * self.pitch
        self.dH = torch.empty(cells, dtype=torch.int32, device=dev)
        self.dP = torch.empty(cells, dtype=torch.int32, device=dev)
        self.d_pos = torch.zeros(1, dtype=torch.int64, device=dev)
        self.d_score = torch.zeros(1, dtype=torch.int32, device=dev)
        self.d_len = torch.zeros(1, dtype=torch.int64, device=dev)
        self.d_end = torch.zeros(1, dtype=torch.int64, device=dev)
        self.nflags = int(lib.swb_strip_flag_count(self.n))
        # boundary buffers (cudaMalloc: they may be mapped by another process), double-buffered by epoch parity
        # so that a left neighbour that is one call ahead never overwrites values still being read
        self.left_in, self.left_flags = [0, 0], [0, 0]
        if not first:
            for k in range(2):
                self.left_in[k] = lib.swb_ipc_alloc(4 * (self.n + 1), device)
                self.left_flags[k] = lib.swb_ipc_alloc(4 * max(self.nflags, 1), device)
                if not self.left_in[k] or not self.left_flags[k]:
                    raise SwbError("boundary buffer allocation failed")
        self.right_out, self.right_flags = [0, 0], [0, 0]
        self._opened = []
        self.epoch = 0

    # -- wiring ---------------------------------------------------------------------------------
    def handles(self) -> bytes:
        """IPC handles of my boundary buffers (4 x 64 bytes), for the strip on my left."""
        out = bytearray(256)
        if self.first:
            return bytes(out)
        buf = (C.c_ubyte * 64)()
        for k, p in enumerate([self.left_in[0], self.left_in[1], self.left_flags[0], self.left_flags[1]]):
            _check(lib.swb_ipc_get_handle(p, buf))
            out[64 * k:64 * k + 64] = bytes(buf)
        return bytes(out)

    def connect_right_ipc(self, handles: bytes) -> None:
        """Map the boundary buffers of the strip on my right (another process)."""
        ptrs = []
        for k in range(4):
            h = (C.c_ubyte * 64).from_buffer_copy(handles[64 * k:64 * k + 64])
            p = C.c_void_p(0)
            _check(lib.swb_ipc_open(h, self.device, C.byref(p)))
            ptrs.append(p.value)
            self._opened.append(p.value)
        self.right_out, self.right_flags = ptrs[0:2], ptrs[2:4]

    def connect_right_local(self, right: "ColumnStrip") -> None:
        """Same process: the right strip's buffers are directly addressable (peer access across devices)."""
        if right.device != self.device:
            _check(lib.swb_enable_peer(self.device, right.device))
        self.right_out, self.right_flags = list(right.left_in), list(right.left_flags)

    # -- compute --------------------------------------------------------------------------------
    def fill_async(self, stream=None, timer=None) -> None:
        self.epoch += 1
        k = self.epoch & 1
        tun = Tuning(timer=timer._h if timer is not None else None)
        _check(lib.swb_fill_strip_async(_ptr(self.a_d), self.m, _ptr(self.b_d), self.n, C.byref(self.scoring),
                                        _ptr(self.dH), _ptr(self.dP), self.pitch,
                                        self.left_in[k] or None, self.left_flags[k] or None,
                                        self.right_out[k] or None, self.right_flags[k] or None, self.epoch,
                                        _ptr(self.d_pos), _ptr(self.d_score), self.device, _stream_ptr(stream),
                                        C.byref(tun)))

    def local_max(self):
        """-> (score, i, j_global) of the local maximum (synchronises)."""
        score = int(self.d_score.item())
        pos = int(self.d_pos.item())
        if score <= 0:
            return (0, 0, 0)
        i, jl = divmod(pos, self.pitch)
        return (score, i, self.col0 + jl)

    def walk(self, i: int, j_local: int, stream=None):
        """backtrack from my cell (i, j_local) -> (path_len, end_i, end_j_local)"""
        _check(lib.swb_backtrack_from_async(_ptr(self.dP), self.pitch, i * self.pitch + j_local, _ptr(self.d_len),
                                            _ptr(self.d_end), self.device, _stream_ptr(stream)))
        plen, end = int(self.d_len.item()), int(self.d_end.item())
        ei, ej = divmod(end, self.pitch)
        return plen, ei, ej

    def matrices(self):
        """(H, P) of my OWN columns as (n+1) x m_local torch tensors (views; local column 0 dropped)."""
        H = self.dH.view(self.n + 1, self.pitch)[:, 1:]
        P = self.dP.view(self.n + 1, self.pitch)[:, 1:]
        return H, P

    def close(self) -> None:
        for p in self._opened:
            lib.swb_ipc_close(p, self.device)
        self._opened = []
        if not self.first:
            for k in range(2):
                if self.left_in[k]:
                    lib.swb_ipc_free(self.left_in[k], self.device)
                if self.left_flags[k]:
                    lib.swb_ipc_free(self.left_flags[k], self.device)
            self.left_in, self.left_flags = [0, 0], [0, 0]
        # the matrices go back to the allocator with the strip (tens of GB at the full-size configurations)
        self.dH = self.dP = self.a_d = self.b_d = None


# ----------------------------------------------------------------------------------------------
# driver 1: one process, several strips (same device or one device per strip)
# ----------------------------------------------------------------------------------------------
class StripSet:
    def __init__(self, a: bytes, b: bytes, nstrips: int, devices=None, scoring=None):
        import torch
        self.torch = torch
        self.m, self.n = len(a), len(b)
        self.parts = partition(self.m, nstrips)
        devices = list(devices) if devices is not None else [0] * nstrips
        self.strips = [ColumnStrip(a[c0:c0 + w], b, c0, self.m, devices[g], g == 0, scoring)
                       for g, (c0, w) in enumerate(self.parts)]
        for g in range(nstrips - 1):
            self.strips[g].connect_right_local(self.strips[g + 1])
        self.concurrent = len(set(devices)) == nstrips
        self.streams = [torch.cuda.Stream(device=d) for d in devices] if self.concurrent else None

    def fill(self):
        """-> maxPos (index into the (n+1) x (m+1) matrix of the whole pair)"""
        torch = self.torch
        if self.concurrent:
            # one device per strip: all strips run at once, each a few row-blocks behind its left neighbour
            for s, st in zip(self.strips, self.streams):
                with torch.cuda.device(s.device):
                    s.fill_async(stream=st)
            for s in self.strips:
                torch.cuda.synchronize(s.device)
        else:
            # shared device: a strip needs its left neighbour's boundary, so they run left to right in stream order
            for s in self.strips:
                with torch.cuda.device(s.device):
                    s.fill_async()
            torch.cuda.synchronize()
        return reduce_maxpos([s.local_max() for s in self.strips], self.m)

    def backtrack(self, maxPos: int):
        return chain_backtrack(self.parts, maxPos, self.m, lambda g, i, jl: self.strips[g].walk(i, jl))

    def gather(self):
        """whole-pair (H, P) as numpy (n+1) x (m+1) (tests only)"""
        import numpy as np
        H = np.zeros((self.n + 1, self.m + 1), np.int32)
        P = np.zeros((self.n + 1, self.m + 1), np.int32)
        for s in self.strips:
            h, p = s.matrices()
            H[:, s.col0 + 1:s.col0 + 1 + s.m] = h.cpu().numpy()
            P[:, s.col0 + 1:s.col0 + 1 + s.m] = p.cpu().numpy()
        return H, P

    def close(self):
        for s in self.strips:
            s.close()


# ----------------------------------------------------------------------------------------------
# driver 2: one process per GPU (torch.distributed)
# ----------------------------------------------------------------------------------------------
def allgather_maxpos(local, m: int, dist, device=None) -> int:
    """all-gather of (score, i, j_global) over the ranks + the reference's tie-break; works on gloo (CPU) and NCCL."""
    import torch
    t = torch.tensor(list(local), dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return reduce_maxpos([tuple(int(x) for x in o.tolist()) for o in out], m)


def distributed_backtrack(parts, maxPos: int, m: int, rank: int, walk_local, dist, device=None):
    """Every rank calls this; rank g runs walk_local(i, j_local) when the path is in its strip and broadcasts
    (path_len, end_i, end_j_local).  -> total path length (same on every rank)."""
    import torch

    def walk(g, i, jl):
        t = torch.zeros(3, dtype=torch.int64, device=device)
        if g == rank:
            plen, ei, ej = walk_local(i, jl)
            t = torch.tensor([plen, ei, ej], dtype=torch.int64, device=device)
        dist.broadcast(t, src=g)
        return tuple(int(x) for x in t.tolist())

    total, _ = chain_backtrack(parts, maxPos, m, walk)
    return total


class StripPipeline:
    """One rank's strip of a pair; torch.distributed must be initialised (one process per GPU)."""

    def __init__(self, a: bytes, b: bytes, device: int, scoring=None, dist=None):
        import torch
        import torch.distributed as tdist
        self.torch, self.dist = torch, dist or tdist
        self.rank, self.world = self.dist.get_rank(), self.dist.get_world_size()
        self.m, self.n = len(a), len(b)
        self.parts = partition(self.m, self.world)
        c0, w = self.parts[self.rank]
        self.device = device
        self.strip = ColumnStrip(a[c0:c0 + w], b, c0, self.m, device, self.rank == 0, scoring)
        self.tdev = torch.device("cuda", device)
        # boundary buffers: every rank publishes the IPC handles of its own, rank g maps those of rank g+1
        mine = torch.frombuffer(bytearray(self.strip.handles()), dtype=torch.uint8).to(self.tdev)
        allh = [torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allh, mine)
        if self.rank + 1 < self.world:
            self.strip.connect_right_ipc(bytes(allh[self.rank + 1].cpu().numpy().tobytes()))
        self.dist.barrier()
        self._unjoined_fills = 0

    def fill_async(self, stream=None, timer=None):
        # Back-pressure (the boundary buffers are double-buffered by call parity): a rank may start call k+1 only after
        # every rank has finished call k-1, or the left GPU would overwrite values its right neighbour is still
        # reading.  maxpos() is such a point (local_max() synchronises this rank's fill, the all-gather joins the
        # ranks); two fills without it in between are separated by an explicit barrier here.
        if self._unjoined_fills >= 1:
            self.torch.cuda.synchronize(self.device)
            self.dist.barrier()
            self._unjoined_fills = 0
        self.strip.fill_async(stream=stream, timer=timer)
        self._unjoined_fills += 1

    def maxpos(self) -> int:
        mp = allgather_maxpos(self.strip.local_max(), self.m, self.dist, self.tdev)
        self._unjoined_fills = 0
        return mp

    def backtrack(self, maxPos: int, stream=None) -> int:
        return distributed_backtrack(self.parts, maxPos, self.m, self.rank,
                                     lambda i, jl: self.strip.walk(i, jl, stream=stream), self.dist, self.tdev)

    def close(self):
        self.dist.barrier()
        self.strip.close()
