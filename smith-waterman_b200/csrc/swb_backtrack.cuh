// swb_backtrack.cuh -- backtrack (omp_smithW.c:405-420) with jump tables: follow P from the start cell until a NONE
// cell (or the hand-off marker of a column strip), negate the path in place, optionally emit the moves.
//
// The walk is a serial pointer chase; a single thread pays ~68 clk per cell even from shared memory (the first
// kernel of this file's predecessor, swb::backtrack_kernel, did exactly that: 2.05 ms for the 53 856 cells of the
// 45000 x 45000 pair).  Here the chase is shortened instead of sped up:
//   * P is staged in diagonal BANDS of 128 rows x 128 columns (band row rr holds the 128 columns around
//     c0 - (i0 - r): the path runs up-left along a diagonal with small drift), three band buffers;
//   * HELPER warps (15 of the 16) turn every staged band into a JUMP TABLE by pointer doubling: the word of a cell
//     becomes (P code, cells skipped, cell reached), three rounds in place -- a jump then covers 8..32 cells.  Only
//     the 64 band columns around the column the path entered the previous band at are doubled (a band follows the
//     diagonal, so the path keeps its band column up to its drift); every other cell keeps its single-move word,
//     which is just as valid: a path that strays outside the window walks there cell by cell;
//   * the WALKER (one thread) follows jumps through the current band: ~25 dependent shared-memory loads per band
//     instead of ~190, and records (first cell, count) per jump;
//   * while it walks, the helpers negate the previous band's path in global memory (each recorded jump is replayed
//     cell by cell by one thread, from the P codes the table keeps), build the table of the band that continues the
//     diagonal (its rows were fetched one band earlier with one TMA bulk copy per row) and start fetching the one
//     after that.  If the walk leaves a band sideways or enters the next one too far off centre, a band centred on
//     the current cell is fetched instead.
// The first and last column of a band and the row above it hold the marker 4: stepping on it ends the walk in this
// band.  Optional output: the moves of the path in walk order (from the start cell backwards), one byte per cell
// (1 UP, 2 LEFT, 3 DIAGONAL) -- the raw material of a CIGAR string.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace swb {
namespace bt2 {

constexpr int kRows = 128;
constexpr int kCols = 128;
constexpr int kRS = 132;                           // staged ints per band row: 33 aligned 16-byte chunks
constexpr int kThreads = 512;
constexpr int kHelpers = kThreads - 32;
constexpr int kBandInts = kRows * kRS;
constexpr int kPad = 160;                          // ints before each band = the marker row above it
constexpr int kBufInts = kPad + kBandInts;
constexpr int kBufs = 3;
constexpr int kList = 256;                         // a band holds at most 128 + 126 path cells
constexpr int kMark = 4;
constexpr int kRounds = 3;                         // pointer-doubling rounds: jumps of >= 8 cells
constexpr int kMaxJump = 32;                       // ... and at most this many (bounds the replay of one jump)
constexpr int kWin = 64;                           // band columns whose words are doubled (power of two)
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

// helpers: start the fetch of a band (one TMA bulk copy per row, counted by the buffer's mbarrier)
__device__ __forceinline__ void fetch_issue(int* buf, const int32_t* P, long long pitch, long long limit,
                                            long long i0, long long c0, unsigned long long* mbar)
{
    const int tid = threadIdx.x - 32;
    const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
    if (tid < kRows) {
        const int rr = tid;
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

__device__ __forceinline__ void helper_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kHelpers) : "memory"); }

// helpers: wait for the fetch, shift every row into place, build the jump table
__device__ __forceinline__ void fetch_finish_and_build(int* buf, long long pitch, long long i0, long long c0,
                                                       unsigned long long* mbar, unsigned& phase, const int centre)
{
    const int tid = threadIdx.x - 32;
    const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
    {
        unsigned done = 0;
        while (!done) {
            asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                         : "=r"(done) : "r"(mb), "r"(phase) : "memory");
        }
        phase ^= 1u;
    }
    helper_bar();
    // A row landed with its first band column at offset (row start & 3): shift it into place (one warp per row, in
    // place: all loads of the row precede its stores) and turn the P values into single-move table words.
    {
        const int lane = tid & 31, hw = tid >> 5;
        for (int rr = hw; rr < kRows; rr += kHelpers / 32) {
            long long r;
            const int sh = (int)(row_start(i0, c0, pitch, rr, r) & 3);
            int* row = buf + rr * kRS;
            int v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = row[sh + 4 * lane + e];
            __syncwarp();
            if (lane == 0) v[0] = kMark;                            // marker columns (band columns 0 and 127)
            if (lane == 31) v[3] = kMark;
            int w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int a = rr * kRS + 4 * lane + e;
                const int code = v[e] < 0 ? 6 : (v[e] & 7);
                const bool move = code >= 1 && code <= 3;
                w[e] = rec_make(code, move ? 1 : 0, move ? a - move_step(code) : a);
            }
            *reinterpret_cast<int4*>(row + 4 * lane) = make_int4(w[0], w[1], w[2], w[3]);
            // (the four staged ints past band column 127 are never reached; clear them all the same)
            if (lane == 0) *reinterpret_cast<int4*>(row + kCols) = make_int4(0, 0, 0, 0);
        }
    }
    helper_bar();
    // pointer doubling, in place, over the window of kWin band columns around `centre`.  A word is a single 32-bit
    // store and every value it ever holds is a valid jump (the cell reached after `cnt` moves), so concurrent
    // readers may see the old or the new word: both are right.  Four cells per thread and pass: loads first.
    const int clo = min(max(centre - kWin / 2, 0), kCols - kWin);
    for (int round = 0; round < kRounds; ++round) {
        for (int base = 4 * tid; base < kRows * kWin; base += 4 * kHelpers) {
            int a[4], r[4], t[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int idx = base + e;
                a[e] = (idx / kWin) * kRS + clo + (idx & (kWin - 1));
                r[e] = buf[a[e]];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) t[e] = buf[rec_tgt(r[e])];           // (a terminal word points at itself)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c2 = rec_cnt(r[e]) + rec_cnt(t[e]);
                if (rec_cnt(r[e]) != 0 && rec_cnt(t[e]) != 0 && c2 <= kMaxJump)
                    buf[a[e]] = rec_make(rec_code(r[e]), c2, rec_tgt(t[e]));
            }
        }
        helper_bar();
    }
}

// helpers: negate the cells of the recorded jumps of a finished band in global memory (*= PATH, :417) and emit the
// moves.  list entry = first cell (15 bits, + kPad) | count << 16 | offset in the band's path << 22
__device__ __forceinline__ void writeback(const int* buf, const int* list, int nlist, int32_t* P, long long pitch,
                                          long long i0, long long c0, unsigned char* ops, long long ops_base)
{
    const int tid = threadIdx.x - 32;
    for (int e = tid; e < nlist; e += kHelpers) {
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
    helper_bar();                                                   // all of it read before the buffer is reused
}

__global__ void __launch_bounds__(kThreads)
backtrack_kernel(int32_t* P, long long pitch, long long maxPos_arg, const long long* d_maxPos,
                 long long* d_pathLen, long long* d_endPos, unsigned char* d_ops)
{
    extern __shared__ __align__(16) int bt_smem[];      // kBufs x (pad + 128x132 ints), then kBufs lists
    __shared__ long long s_len;
    __shared__ int s_done, s_a, s_count[kBufs];
    __shared__ __align__(8) unsigned long long s_mbar[kBufs];     // one per band buffer: counts the rows of a fetch
    unsigned mphase[kBufs] = {0u, 0u, 0u};
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
    const bool walker = threadIdx.x == 0, helper = threadIdx.x >= 32;
    auto bandbuf = [&](int b) { return bt_smem + b * kBufInts + kPad; };
    auto listbuf = [&](int b) { return bt_smem + kBufs * kBufInts + b * kList; };
    if (walker) { s_len = 0; s_done = 0; for (int k = 0; k < kBufs; ++k) s_count[k] = 0; }
    // the marker row above every band (terminal table words that point at themselves)
    for (int e = threadIdx.x; e < kBufs * kPad; e += kThreads) {
        const int b = e / kPad, x = e % kPad;
        bt_smem[b * kBufInts + x] = rec_make(kMark, 0, x - kPad);
    }
    // buffers: cur = the band being walked, nxt = the band that continues its diagonal (table built while cur is
    // walked), third = the previous band (written back first) and then the band after nxt (fetch in flight)
    int cur = 0, nxt = 1, third = 2;
    long long bi = i, bc = j;                           // entry cell of the band in `cur`
    long long pbi = 0, pbc = 0, plen0 = 0;              // previous band (in `third`): entry cell, path offset of its first cell
    bool have_prev = false;
    __syncthreads();
    if (helper) {
        fetch_issue(bandbuf(cur), P, pitch, limit, bi, bc, &s_mbar[cur]);
        fetch_finish_and_build(bandbuf(cur), pitch, bi, bc, &s_mbar[cur], mphase[cur], kCols / 2);
        fetch_issue(bandbuf(nxt), P, pitch, limit, bi - kRows, bc - kRows, &s_mbar[nxt]);
    }
    __syncthreads();
    int k0 = kCols / 2;                                 // band column of the current cell (bottom row)
    long long len_before = 0;                           // path cells before the band in `cur`
    while (true) {
        const long long ni = bi - kRows, nc = bc - kRows;          // the band that continues this one's diagonal
        if (helper) {
            if (have_prev) writeback(bandbuf(third), listbuf(third), s_count[third], P, pitch, pbi, pbc, d_ops, plen0);
            // the band after next goes into the buffer that was just written back; then finish the next band
            fetch_issue(bandbuf(third), P, pitch, limit, ni - kRows, nc - kRows, &s_mbar[third]);
            fetch_finish_and_build(bandbuf(nxt), pitch, ni, nc, &s_mbar[nxt], mphase[nxt], k0);
        } else if (walker) {
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
            // the last band's path
            if (helper) {
                // (the fetch issued into `third` is still in flight: wait for it before the kernel ends)
                const unsigned mb = (unsigned)__cvta_generic_to_shared(&s_mbar[third]);
                unsigned done = 0;
                while (!done) {
                    asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                                 : "=r"(done) : "r"(mb), "r"(mphase[third]) : "memory");
                }
                writeback(bandbuf(cur), listbuf(cur), s_count[cur], P, pitch, bi, bc, d_ops, len_cur);
            }
            break;
        }
        // usable prefetch: the walk left through the top and enters the next band well inside it
        const long long kn = j - (nc - kCols / 2);                  // its column in the next band's bottom row
        if (i == ni && kn >= 24 && kn <= kCols - 24) {
            // rotate: cur -> previous (third), nxt -> cur, third (fetch in flight) -> nxt
            pbi = bi; pbc = bc; plen0 = len_cur; have_prev = true;
            const int old_cur = cur;
            cur = nxt; nxt = third; third = old_cur;
            bi = ni; bc = nc;
            k0 = (int)kn;
            __syncthreads();                                         // s_* read by everyone before the walker rewrites them
        } else {
            // the prefetched bands are useless: write back the finished one now and fetch a band centred here
            __syncthreads();
            if (helper) {
                // drain the fetch in flight in `third` (its buffer is reused below), then restart the pipeline
                const unsigned mb = (unsigned)__cvta_generic_to_shared(&s_mbar[third]);
                unsigned done = 0;
                while (!done) {
                    asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                                 : "=r"(done) : "r"(mb), "r"(mphase[third]) : "memory");
                }
                mphase[third] ^= 1u;
                helper_bar();
                writeback(bandbuf(cur), listbuf(cur), s_count[cur], P, pitch, bi, bc, d_ops, len_cur);
                fetch_issue(bandbuf(cur), P, pitch, limit, i, j, &s_mbar[cur]);
                fetch_finish_and_build(bandbuf(cur), pitch, i, j, &s_mbar[cur], mphase[cur], kCols / 2);
                fetch_issue(bandbuf(nxt), P, pitch, limit, i - kRows, j - kRows, &s_mbar[nxt]);
            }
            have_prev = false;
            bi = i; bc = j;
            k0 = kCols / 2;
            __syncthreads();
        }
    }
    if (threadIdx.x == 0 && d_pathLen) *d_pathLen = s_len;
    if (threadIdx.x == 0 && d_endPos) *d_endPos = i * pitch + j;
}

constexpr int kSmemBytes = (kBufs * kBufInts + kBufs * kList) * (int)sizeof(int);

}  // namespace bt2
}  // namespace swb
