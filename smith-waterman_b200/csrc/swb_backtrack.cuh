// swb_backtrack.cuh -- backtrack (omp_smithW.c:405-420) with jump tables: follow P from the start cell until a NONE
// cell (or the hand-off marker of a column strip), negate the path in place, optionally emit the moves.
//
// The walk is a serial pointer chase; a single thread pays ~68 clk per cell even from shared memory (the first
// backtrack kernel, swb::backtrack_kernel in swb_kernels.cuh, does exactly that: 2.05 ms for the 53 856 cells of the
// 45000 x 45000 pair).  Here the chase is shortened instead of sped up:
//   * P is staged in diagonal BANDS of 128 rows x 128 columns (band row rr holds the 128 columns around
//     c0 - (i0 - r): the path runs up-left along a diagonal with small drift), three band buffers;
//   * BUILDER warps turn every staged band into a JUMP TABLE by pointer doubling: the word of a cell becomes
//     (P code, cells skipped, cell reached), SWB_BT_ROUNDS rounds in place (jumps of 2^rounds cells).  Only the
//     kWin band columns around the column the path entered the current band at are doubled (a band follows the
//     diagonal, so the path keeps its band column up to its drift); every other cell keeps its single-move word, which
//     is just as valid: a path that strays outside the window walks there cell by cell;
//   * the WALKER (one thread) follows jumps through the current band: a few dozen dependent shared-memory loads per
//     band instead of ~190, and records (first cell, count) per jump;
//   * meanwhile WRITE-BACK warps negate the previous band's path in global memory (each recorded jump is replayed cell
//     by cell by one thread, from the P codes the table keeps) and start the fetch (one TMA bulk copy per row) of the
//     band after the next one, centred on where the path is heading.  If the walk leaves a band sideways or enters the
//     next one too far off centre, a band centred on the current cell is fetched instead.
// All shared-memory sweeps are laid out lane-consecutive: the table build is bound by shared-memory wavefronts (a
// stride-4 layout made it 4x slower than the walk it was meant to shorten).
// Measured on the 53 856-cell path of the 45000 x 45000 pair (gpurun_out/r02g): 0 rounds 4.9 M clk, 1 round 4.0 M,
// 2 rounds 4.6 M, 3 rounds 5.5 M -- building the table for 8192 window cells costs more than it saves on a path that
// visits ~150 of them, so ONE round is the default and the kernel only equals the single-walker kernel of
// swb_kernels.cuh (4.0 M clk = 2.05 ms).  It is used where the moves of the path are wanted (swb_traceback_async).
// The first and last column of a band and the row above it hold the marker 4: stepping on it ends the walk in this
// band.  Optional output: the moves of the path in walk order (from the start cell backwards), one byte per cell
// (1 UP, 2 LEFT, 3 DIAGONAL) -- the raw material of a CIGAR string.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace swb {
namespace bt2 {

constexpr int kRows = 128;
constexpr int kCols = 128;
constexpr int kRS = 132;                           // staged ints per band row: 33 aligned 16-byte chunks
constexpr int kThreads = 512;
constexpr int kBuilders = 256;                     // warps 1..8
constexpr int kWriters = kThreads - 32 - kBuilders;    // warps 9..15 (>= kRows: one fetch thread per band row)
constexpr int kBandInts = kRows * kRS;
constexpr int kPad = 160;                          // ints before each band = the marker row above it
constexpr int kBufInts = kPad + kBandInts;
constexpr int kBufs = 3;
constexpr int kList = 256;                         // a band holds at most 128 + 126 path cells
constexpr int kMark = 4;
#ifndef SWB_BT_ROUNDS
#define SWB_BT_ROUNDS 1
#endif
#ifndef SWB_BT_WIN
#define SWB_BT_WIN 64
#endif
constexpr int kRounds = SWB_BT_ROUNDS;             // pointer-doubling rounds: jumps of >= 2^rounds cells ...
constexpr int kMaxJump = 16;                       // ... and at most this many (bounds the replay of one jump)
constexpr int kWin = SWB_BT_WIN;                   // band columns whose words are doubled (power of two)
constexpr int kEnterLo = 24, kEnterHi = kCols - 24;    // a prefetched band is usable if the path enters it in these columns
static_assert(kWriters >= kRows, "one fetch thread per band row");
// table word: bits 0-2 raw code (P 0..5, 6 = already negated), bits 3-8 cells skipped, bits 12-26 index of the cell
// reached (+ kPad: the marker row above the band has negative indices)
__device__ __forceinline__ int rec_code(int r) { return r & 7; }
__device__ __forceinline__ int rec_cnt(int r) { return (r >> 3) & 63; }
__device__ __forceinline__ int rec_tgt(int r) { return (r >> 12) - kPad; }
__device__ __forceinline__ int rec_make(int code, int cnt, int tgt) { return code | (cnt << 3) | ((tgt + kPad) << 12); }
// step of one move inside a staged band: UP (:412) = one band row up and one column right, LEFT (:414) = one column
// left, DIAGONAL (:410) = one band row up (the band follows the diagonal)
__device__ __forceinline__ int move_step(int code) { return code == 1 ? kRS - 1 : code == 2 ? 1 : kRS; }

// global index of band column 0 of band row rr, for the band entered at (i0, c0) (its cell (127, 64))
__device__ __forceinline__ long long row_start(long long i0, long long c0, long long pitch, int rr, long long& r)
{
    r = i0 - (kRows - 1) + rr;
    return r * pitch + (c0 - (i0 - r) - kCols / 2);
}

__device__ __forceinline__ void builder_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kBuilders) : "memory"); }
__device__ __forceinline__ void writer_bar() { asm volatile("bar.sync 2, %0;" ::"n"(kWriters) : "memory"); }

__device__ __forceinline__ void mbar_wait(unsigned long long* mbar, unsigned& phase)
{
    const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
    unsigned done = 0;
    while (!done) {
        asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                     : "=r"(done) : "r"(mb), "r"(phase) : "memory");
    }
    phase ^= 1u;
}

// write-back warps: start the fetch of a band (one TMA bulk copy per row, counted by the buffer's mbarrier)
__device__ __forceinline__ void fetch_issue(int wtid, int* buf, const int32_t* P, long long pitch, long long limit,
                                            long long i0, long long c0, unsigned long long* mbar)
{
    const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
    if (wtid < kRows) {
        const int rr = wtid;
        long long r;
        const long long s = row_start(i0, c0, pitch, rr, r);
        const long long g0 = s & ~3LL;                             // aligned global index of the row's first chunk
        int* dst = buf + rr * kRS;
        if (r >= 0 && g0 >= 0 && g0 + kRS <= limit) {
            const unsigned sd = (unsigned)__cvta_generic_to_shared(dst);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the buffer was last touched through the generic proxy
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(kRS * 4) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sd), "l"(P + g0), "r"(kRS * 4), "r"(mb) : "memory");
        } else {
            for (int k = 0; k < kRS; ++k) dst[k] = (r >= 0 && g0 + k >= 0 && g0 + k < limit) ? P[g0 + k] : 0;
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
        }
    }
}

// builder warps: wait for the fetch, shift every row into place, build the jump table
__device__ __forceinline__ void build(int btid, int* buf, long long pitch, long long i0, long long c0,
                                      unsigned long long* mbar, unsigned& phase, const int centre)
{
    mbar_wait(mbar, phase);
    builder_bar();
    // A row landed with its first band column at offset (row start & 3): shift it into place (one warp per row, in
    // place: all loads of the row precede its stores) and turn the P values into single-move table words.  Lane l
    // holds band columns l, l+32, l+64, l+96 (consecutive lanes = consecutive words: no bank conflicts).
    {
        const int lane = btid & 31, hw = btid >> 5;
        for (int rr = hw; rr < kRows; rr += kBuilders / 32) {
            long long r;
            const int sh = (int)(row_start(i0, c0, pitch, rr, r) & 3);
            int* row = buf + rr * kRS;
            int v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = row[sh + lane + 32 * e];
            __syncwarp();
            if (lane == 0) v[0] = kMark;                            // marker columns (band columns 0 and 127)
            if (lane == 31) v[3] = kMark;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int a = rr * kRS + lane + 32 * e;
                const int code = v[e] < 0 ? 6 : (v[e] & 7);
                const bool move = code >= 1 && code <= 3;
                row[lane + 32 * e] = rec_make(code, move ? 1 : 0, move ? a - move_step(code) : a);
            }
            // (the four staged ints past band column 127 are never reached; clear them all the same)
            if (lane < 4) row[kCols + lane] = 0;
        }
    }
    builder_bar();
    // pointer doubling, in place, over the window of kWin band columns around `centre`.  A word is a single 32-bit
    // store and every value it ever holds is a valid jump (the cell reached after `cnt` moves), so concurrent
    // readers may see the old or the new word: both are right.  Four cells per thread and pass, a warp covering
    // half a window row per load (consecutive lanes = consecutive words): loads first, then the dependent loads.
    const int clo = min(max(centre - kWin / 2, 0), kCols - kWin);
    for (int round = 0; round < kRounds; ++round) {
        for (int base = btid; base < kRows * kWin; base += 4 * kBuilders) {
            int a[4], r[4], t[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int idx = base + e * kBuilders;
                a[e] = (idx / kWin) * kRS + clo + (idx & (kWin - 1));
                r[e] = buf[a[e]];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) t[e] = buf[rec_tgt(r[e])];                  // (a terminal word points at itself)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c2 = rec_cnt(r[e]) + rec_cnt(t[e]);
                if (rec_cnt(r[e]) != 0 && rec_cnt(t[e]) != 0 && c2 <= kMaxJump)
                    buf[a[e]] = rec_make(rec_code(r[e]), c2, rec_tgt(t[e]));
            }
        }
        builder_bar();
    }
}

// write-back warps: negate the cells of the recorded jumps of a finished band in global memory (*= PATH, :417) and
// emit the moves.  list entry = first cell (+ kPad, 16 bits) | count << 16 | offset in the band's path << 22
__device__ __forceinline__ void writeback(int wtid, const int* buf, const int* list, int nlist, int32_t* P, long long pitch,
                                          long long i0, long long c0, unsigned char* ops, long long ops_base)
{
    for (int e = wtid; e < nlist; e += kWriters) {
        const int ent = list[e];
        int a = (ent & 0xffff) - kPad;
        const int cnt = (ent >> 16) & 63;
        const int off = (ent >> 22) & 1023;
        for (int k = 0; k < cnt; ++k) {
            const int code = rec_code(buf[a]);
            const int rr = a / kRS;
            long long r;
            const long long s = row_start(i0, c0, pitch, rr, r);
            P[s + (a - rr * kRS)] = -code;
            if (ops) ops[ops_base + off + k] = (unsigned char)code;
            a -= move_step(code);
        }
    }
    writer_bar();                                                   // all of it read before the buffer is reused
}

__global__ void __launch_bounds__(kThreads)
backtrack_kernel(int32_t* P, long long pitch, long long maxPos_arg, const long long* d_maxPos,
                 long long* d_pathLen, long long* d_endPos, unsigned char* d_ops)
{
    extern __shared__ __align__(16) int bt_smem[];      // kBufs x (pad + 128x132 ints), then kBufs lists
    __shared__ long long s_len;
    __shared__ int s_done, s_a, s_count[kBufs];
    __shared__ __align__(8) unsigned long long s_mbar[kBufs];     // one per band buffer: counts the rows of a fetch
    unsigned mphase[kBufs] = {0u, 0u, 0u};              // (tracked by the builders, who do all the waiting)
    if (threadIdx.x == 0) {
        for (int k = 0; k < kBufs; ++k)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_mbar[k])), "r"(kRows) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long pos = d_maxPos ? *d_maxPos : maxPos_arg;
    if (pos <= 0) { if (threadIdx.x == 0) { if (d_pathLen) *d_pathLen = 0; if (d_endPos) *d_endPos = 0; } return; }
    long long i = pos / pitch, j = pos % pitch;         // current cell
    // the path only moves up and left: nothing after the end of the start row is ever needed (or read)
    const long long limit = (i + 1) * pitch;
    const bool walker = threadIdx.x == 0;
    const bool builder = threadIdx.x >= 32 && threadIdx.x < 32 + kBuilders;
    const bool writer = threadIdx.x >= 32 + kBuilders;
    const int btid = threadIdx.x - 32, wtid = threadIdx.x - 32 - kBuilders;
    auto bandbuf = [&](int b) { return bt_smem + b * kBufInts + kPad; };
    auto listbuf = [&](int b) { return bt_smem + kBufs * kBufInts + b * kList; };
    if (walker) { s_len = 0; s_done = 0; for (int k = 0; k < kBufs; ++k) s_count[k] = 0; }
    // the marker row above every band (terminal table words that point at themselves)
    for (int e = threadIdx.x; e < kBufs * kPad; e += kThreads) {
        const int b = e / kPad, x = e % kPad;
        bt_smem[b * kBufInts + x] = rec_make(kMark, 0, x - kPad);
    }
    // buffers: cur = the band being walked, nxt = the band the path is expected to enter next (its table is built while
    // cur is walked), third = the previous band (written back first), then the band after nxt (fetch in flight)
    int cur = 0, nxt = 1, third = 2;
    long long bi = i, bc = j;                           // entry cell of the band in `cur` (its cell (127, 64))
    long long ni = i - kRows, nc = j - kRows;           // ... of the band in `nxt`
    long long pbi = 0, pbc = 0, plen0 = 0;              // previous band (in `third`): entry cell, path offset of its first cell
    bool have_prev = false;
    __syncthreads();
    if (writer) {
        fetch_issue(wtid, bandbuf(cur), P, pitch, limit, bi, bc, &s_mbar[cur]);
        fetch_issue(wtid, bandbuf(nxt), P, pitch, limit, ni, nc, &s_mbar[nxt]);
    }
    if (builder) build(btid, bandbuf(cur), pitch, bi, bc, &s_mbar[cur], mphase[cur], kCols / 2);
    __syncthreads();
    int k0 = kCols / 2;                                 // band column of the current cell (bottom row)
    long long len_before = 0;                           // path cells before the band in `cur`
#ifdef SWB_X_BTDEBUG
    long long dbg_walk = 0, dbg_wb = 0, dbg_build = 0, dbg_total0 = clock64(); int dbg_bands = 0, dbg_miss = 0, dbg_jumps = 0;
#endif
    while (true) {
        // the band after next: two bands up the diagonal from where the path entered this one
        const long long ti = ni - kRows, tc = bc + (k0 - kCols / 2) - 2 * kRows;
        if (writer) {
#ifdef SWB_X_BTDEBUG
            const long long h0 = clock64();
#endif
            if (have_prev) writeback(wtid, bandbuf(third), listbuf(third), s_count[third], P, pitch, pbi, pbc, d_ops, plen0);
            fetch_issue(wtid, bandbuf(third), P, pitch, limit, ti, tc, &s_mbar[third]);
#ifdef SWB_X_BTDEBUG
            dbg_wb += clock64() - h0;
#endif
        } else if (builder) {
#ifdef SWB_X_BTDEBUG
            const long long h0 = clock64();
#endif
            build(btid, bandbuf(nxt), pitch, ni, nc, &s_mbar[nxt], mphase[nxt], k0);
#ifdef SWB_X_BTDEBUG
            dbg_build += clock64() - h0;
#endif
        } else if (walker) {
#ifdef SWB_X_BTDEBUG
            const long long w0 = clock64(); ++dbg_bands;
#endif
            const int* band = bandbuf(cur);
            int* list = listbuf(cur);
            int a = (kRows - 1) * kRS + k0, n = 0, total = 0;
            int r = band[a];
            while (rec_cnt(r) != 0) {
                list[n++] = (a + kPad) | (rec_cnt(r) << 16) | (total << 22);
                total += rec_cnt(r);
                a = rec_tgt(r);
                r = band[a];
            }
            s_count[cur] = n;
            s_len += total;
            s_done = (rec_code(r) != kMark);                        // NONE, the strip hand-off marker or a negated cell ends the path (:419)
            s_a = a;
#ifdef SWB_X_BTDEBUG
            dbg_walk += clock64() - w0; dbg_jumps += n;
#endif
        }
        __syncthreads();
        const long long len_cur = len_before;
        len_before = s_len;
        {
            const int a = s_a;
            const int rr = (a + kRS) / kRS - 1;                     // -1 = the marker row above the band
            long long r;
            const long long s = row_start(bi, bc, pitch, rr, r);
            const long long g = s + (a - rr * kRS);
            i = r; j = g - r * pitch;
        }
        if (s_done) {
            // the last band's path; the fetch issued into `third` is still in flight: wait for it before the kernel ends
            if (builder) mbar_wait(&s_mbar[third], mphase[third]);
            if (writer) writeback(wtid, bandbuf(cur), listbuf(cur), s_count[cur], P, pitch, bi, bc, d_ops, len_cur);
            break;
        }
        // usable prefetch: the walk left through the top and enters the next band well inside it
        const long long kn = j - (nc - kCols / 2);                  // its column in the next band's bottom row
        if (i == ni && kn >= kEnterLo && kn <= kEnterHi) {
            // rotate: cur -> previous (third), nxt -> cur, third (fetch in flight) -> nxt
            pbi = bi; pbc = bc; plen0 = len_cur; have_prev = true;
            const int old_cur = cur;
            cur = nxt; nxt = third; third = old_cur;
            bi = ni; bc = nc; ni = ti; nc = tc;
            k0 = (int)kn;
            __syncthreads();                                         // s_* read by everyone before the walker rewrites them
        } else {
            // the prefetched bands are useless: write back the finished one, fetch a band centred here and the one
            // that continues its diagonal (rare: the phases are simply separated by block-wide barriers)
#ifdef SWB_X_BTDEBUG
            ++dbg_miss;
#endif
            __syncthreads();
            if (builder) mbar_wait(&s_mbar[third], mphase[third]);  // drain the fetch in flight: its buffer is reused
            if (writer) writeback(wtid, bandbuf(cur), listbuf(cur), s_count[cur], P, pitch, bi, bc, d_ops, len_cur);
            __syncthreads();
            bi = i; bc = j; ni = i - kRows; nc = j - kRows;
            if (writer) {
                fetch_issue(wtid, bandbuf(cur), P, pitch, limit, bi, bc, &s_mbar[cur]);
                fetch_issue(wtid, bandbuf(nxt), P, pitch, limit, ni, nc, &s_mbar[nxt]);
            }
            if (builder) build(btid, bandbuf(cur), pitch, bi, bc, &s_mbar[cur], mphase[cur], kCols / 2);
            have_prev = false;
            k0 = kCols / 2;
            __syncthreads();
        }
    }
    if (threadIdx.x == 0 && d_pathLen) *d_pathLen = s_len;
    if (threadIdx.x == 0 && d_endPos) *d_endPos = i * pitch + j;
#ifdef SWB_X_BTDEBUG
    if (threadIdx.x == 0) printf("bt2 walker: len %lld bands %d misses %d jumps %d walk clk %lld total clk %lld\n", s_len, dbg_bands, dbg_miss, dbg_jumps, dbg_walk, clock64() - dbg_total0);
    if (threadIdx.x == 32) printf("bt2 builder clk %lld\n", dbg_build);
    if (threadIdx.x == 32 + kBuilders) printf("bt2 write-back + fetch issue clk %lld\n", dbg_wb);
#endif
}

constexpr int kSmemBytes = (kBufs * kBufInts + kBufs * kList) * (int)sizeof(int);

}  // namespace bt2
}  // namespace swb
