// swb_kernels.cuh -- sm_100a kernels of the Smith-Waterman fill / backtrack path.
//
// Replaces the nDiag wavefront loop + similarityScore + backtrack of the reference
// (omp_smithW.c:203-216, 331-388, 405-420).  See DESIGN.md for the derivation; the
// short version:
//
//  * The matrix is cut into horizontal STRIPS of 32 rows.  One warp owns a strip and
//    sweeps it left to right; lane l owns row r0+l.  Per STEP every lane computes one
//    16-byte BLOCK (4 consecutive columns) of its row, so H never leaves registers on
//    the dependency chain: the block of the row above arrives by __shfl_up_sync.
//  * pitch = m+1 is in general not a multiple of 4, so "4 consecutive columns" is
//    chosen PER ROW such that every block is a 16-byte aligned int4 of the caller's
//    row-major H/P:  row r has phase phi_r = (r*pitch)&3 and its block qb covers
//    columns 4*qb - phi_r .. +3.  Consecutive rows differ by MU = pitch&3 columns of
//    phase; lane l lags lane l-1 by one step plus that phase (sigma_l extra steps
//    accumulated), and the two most recent blocks of the upper row (8 registers) always
//    contain the 5 upper/diagonal values a block needs -- at compile-time positions
//    (template parameter MU).
//  * A cell is computed on packed keys K = 16*H + tie, tie in {NONE 8, DIAG 7, UP 5,
//    LEFT 2}: one max over the four candidates reproduces the reference's strict-'>'
//    order DIAGONAL, UP, LEFT (omp_smithW.c:348-378), and P = K&3, H = K>>4.  Three
//    DPX VIADDMNMX per cell.
//  * Finished blocks go to a per-warp shared-memory staging ring; whenever a row has 8
//    blocks (128 contiguous bytes) eight lanes write them out with 16-byte stores, so
//    every global store instruction writes four full 128-byte row segments.
//  * Strip -> strip hand-off (row 32 of a strip feeds row 1 of the next):  inside a CTA
//    ("band" of wpc strips) through a shared-memory ring + progress counters, between
//    bands through H itself in global/L2 + a device-scope progress flag (release /
//    acquire).  Bands are claimed from an atomic ticket in start order, so a waiting
//    band's predecessor is always resident: the whole fill is ONE launch, anti-
//    diagonals are not launches.
//  * Per-row maxima fall out of the lane registers; a second tiny kernel scans only the
//    rows that attain the global maximum to find maxPos with the reference's tie-break
//    (first in anti-diagonal order, bottom-left to top-right; omp_smithW.c:203-215,384-387).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace swb {

constexpr int kRingBlocks = 64;     // hand-off ring capacity in 16-byte blocks (power of two)
constexpr int kGroup      = 8;      // steps per synchronisation group == staging ring depth
constexpr int kAOff       = 64;     // leading pad words of the shifted copies of a
constexpr int kMaxWarps   = 16;     // strips per band (CTA) upper bound
constexpr int kWarpSmemBlocks = 32 * kGroup + kRingBlocks;   // int4 per warp

// tie codes: larger wins on equal score => NONE > DIAGONAL > UP > LEFT, and code&3 is
// the reference's P value (omp_smithW.c:33-36)
constexpr int kTieNone = 8, kTieDiag = 7, kTieUp = 5, kTieLeft = 2;

struct FillParams {
    const unsigned* a4;      // 4 phase-shifted word copies of a (built by prep_kernel)
    int             a4_stride;
    const unsigned char* b;  // n bytes, device
    int32_t*        H;
    int32_t*        P;
    long long       pitch;   // ints per row (>= m+1)
    long long       m, n;
    int             s_match, s_mismatch;   // 16*score + kTieDiag
    int             g_up, g_left;          // 16*gap + kTieUp / kTieLeft
    int             steps;                 // steps per strip, multiple of kGroup
    int             qbmax;                 // blocks that can hold valid columns: ((m+3)>>2)+1
    int*            ticket;                // band ticket counter
    int*            progress;              // [nbands+1]; [k+1] = blocks of band k's last row visible in H
    int*            row_max;               // [n+1] max H of each row
    int*            gmax;                  // global max H
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int4 ld_cg_int4(const int4* p)
{
    int4 v;
    asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_cs_int4(int4* p, const int4& v)
{
    asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------
// prep: 4 byte-shifted word copies of a, so that a lane whose blocks start at
// columns == -phi (mod 4) reads the 4 characters of a block with ONE aligned 32-bit load.
//   copy s, word kAOff+qb, byte e  =  a[4*qb + s + e - 4]      (0 outside [0,m))
// A lane of phase phi uses copy s = 3-phi: byte e of word qb is a[col-1] for
// col = 4*qb - phi + e   (matchMissmatchScore reads a[j-1], omp_smithW.c:395).
// Also arms the workspace words.
// ---------------------------------------------------------------------------------
__global__ void prep_kernel(const unsigned char* __restrict__ a, long long m,
                            unsigned* __restrict__ a4, int stride,
                            int* progress, int nprogress, int* ticket, int* gmax,
                            unsigned long long* key)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    for (long long w = tid; w < 4LL * stride; w += nth) {
        const int s = (int)(w / stride);
        const long long k = w % stride;
        const long long base = 4 * (k - kAOff) + s - 4;
        unsigned word = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long idx = base + e;
            const unsigned c = (idx >= 0 && idx < m) ? (unsigned)a[idx] : 0u;
            word |= c << (8 * e);
        }
        a4[w] = word;
    }
    for (long long k = tid; k < nprogress; k += nth) progress[k] = (k == 0) ? 0x7fffffff : 0;
    if (tid == 0) { *ticket = 0; *gmax = 0; *key = ~0ull; }
}

// ---------------------------------------------------------------------------------
// The fill kernel.  grid = number of bands, block = 32*wpc threads,
// dynamic smem = wpc * kWarpSmemBlocks * 16 bytes.
// ---------------------------------------------------------------------------------
template <int MU>
struct Strip {
    // ---- per-lane constants
    int lane, cbase;              // first column of the block of step t is 4*t + cbase
    unsigned b4;
    const unsigned* aw;           // aw[t] = the 4 characters of the block of step t
    int sm, sx, gu, gl;
    int m;
    // ---- dependency state (registers)
    int A[4], B[4];               // the two most recent blocks of the row above (clean 16*H)
    int hl;                       // clean 16*H of the last cell of my previous block
    int rmax;                     // running max of K over my row
    unsigned aword;
    // ---- shared memory
    int4* stage;                  // [32 lanes][8 slots]
    int4* ring_in;                // blocks of the row above my strip
    int4* ring_out;               // blocks of my last row, for the next strip
    bool  has_consumer;
    int   sigma31, wrap0;
    // ---- flush role (lane = 8*fk + fe)
    int fk, fe;
    int rowmask;                  // bit c: row r0 + c + 8*fk exists
    long long g0;                 // int4 index of my flush target at t = 0, c = 0
    long long q4;                 // int4 distance between the targets of rows l and l+1
    int fcol0;                    // column of element 0 of my flush block at t = 0, c = 0
    int4* H4; int4* P4;
    int32_t* H; int32_t* P;

    template <bool EDGE>
    __device__ __forceinline__ void step(const int t)
    {
        // ---------------- cells ----------------
        const unsigned x = aword ^ b4;
        const int s0 = (x & 0x000000ffu) ? sx : sm;       // omp_smithW.c:394-399
        const int s1 = (x & 0x0000ff00u) ? sx : sm;
        const int s2 = (x & 0x00ff0000u) ? sx : sm;
        const int s3 = (x & 0xff000000u) ? sx : sm;
        const int W[8] = {B[0], B[1], B[2], B[3], A[0], A[1], A[2], A[3]};
        const int dg = W[3 - MU], u0 = W[4 - MU], u1 = W[5 - MU], u2 = W[6 - MU], u3 = W[7 - MU];
        // K = max(left+gap|LEFT, up+gap|UP, diag+s|DIAG, 0|NONE)     (omp_smithW.c:339-381)
        int k0 = __viaddmax_s32(hl, gl, __viaddmax_s32(u0, gu, __viaddmax_s32(dg, s0, kTieNone)));
        if (EDGE) { const int c = 4 * t + cbase;     if ((unsigned)(c - 1) >= (unsigned)m) k0 = kTieNone; }
        const int h0 = k0 & ~15;
        int k1 = __viaddmax_s32(h0, gl, __viaddmax_s32(u1, gu, __viaddmax_s32(u0, s1, kTieNone)));
        if (EDGE) { const int c = 4 * t + cbase + 1; if ((unsigned)(c - 1) >= (unsigned)m) k1 = kTieNone; }
        const int h1 = k1 & ~15;
        int k2 = __viaddmax_s32(h1, gl, __viaddmax_s32(u2, gu, __viaddmax_s32(u1, s2, kTieNone)));
        if (EDGE) { const int c = 4 * t + cbase + 2; if ((unsigned)(c - 1) >= (unsigned)m) k2 = kTieNone; }
        const int h2 = k2 & ~15;
        int k3 = __viaddmax_s32(h2, gl, __viaddmax_s32(u3, gu, __viaddmax_s32(u2, s3, kTieNone)));
        if (EDGE) { const int c = 4 * t + cbase + 3; if ((unsigned)(c - 1) >= (unsigned)m) k3 = kTieNone; }
        const int h3 = k3 & ~15;
        hl = h3;
        rmax = __vimax3_s32(rmax, k0, k1);
        rmax = __vimax3_s32(rmax, k2, k3);

        // ---------------- stage my block, hand my row to the next strip ----------------
        stage[lane * 8 + ((t + lane) & 7)] = make_int4(k0, k1, k2, k3);
        if (has_consumer && lane == 31) {
            const int qb31 = t - 31 - sigma31;
            if (!EDGE || qb31 >= 0) ring_out[qb31 & (kRingBlocks - 1)] = make_int4(h0, h1, h2, h3);
        }
        // ---------------- pass my block down one lane ----------------
        B[0] = A[0]; B[1] = A[1]; B[2] = A[2]; B[3] = A[3];
        A[0] = __shfl_up_sync(0xffffffffu, h0, 1);
        A[1] = __shfl_up_sync(0xffffffffu, h1, 1);
        A[2] = __shfl_up_sync(0xffffffffu, h2, 1);
        A[3] = __shfl_up_sync(0xffffffffu, h3, 1);
        if (lane == 0) {
            const int4 v = ring_in[(t + 1 + wrap0) & (kRingBlocks - 1)];
            A[0] = v.x; A[1] = v.y; A[2] = v.z; A[3] = v.w;
        }
        aword = __ldg(aw + t + 1);
        __syncwarp();

        // ---------------- write out the rows that completed 8 blocks ----------------
        // rows l == t+1 (mod 8): lane 8*fk+fe writes block (step t-7+fe) of row c + 8*fk
        {
            const int c  = (t + 1) & 7;
            const int lk = c + 8 * fk;
            const int4 kv = stage[lk * 8 + ((2 * (t + 1) + fe) & 7)];
            const int4 hv = make_int4(kv.x >> 4, kv.y >> 4, kv.z >> 4, kv.w >> 4);
            const int4 pv = make_int4(kv.x & 3, kv.y & 3, kv.z & 3, kv.w & 3);
            const long long g = g0 + t + c * q4;
            if (!EDGE) {
                if ((rowmask >> c) & 1) { st_cs_int4(H4 + g, hv); st_cs_int4(P4 + g, pv); }
            } else {
                if ((rowmask >> c) & 1) {
                    const int col = fcol0 + 4 * t - c * (4 + MU);
                    const int hh[4] = {hv.x, hv.y, hv.z, hv.w};
                    const int pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((unsigned)(col + e) <= (unsigned)m) { H[4 * g + e] = hh[e]; P[4 * g + e] = pp[e]; }
                }
            }
        }
        __syncwarp();
    }
};

template <int MU>
__global__ void __launch_bounds__(32 * kMaxWarps)
fill_kernel(const FillParams p)
{
    extern __shared__ int4 smem4[];
    __shared__ int s_band;
    __shared__ int s_avail[kMaxWarps + 1];      // [w]: blocks of strip w-1's last row present in ring w
    __shared__ int s_consumed[kMaxWarps + 1];   // [w]: ring w entries below this index are free

    const int lane = threadIdx.x & 31;
    const int w    = threadIdx.x >> 5;
    const int wpc  = blockDim.x >> 5;

    if (threadIdx.x == 0) s_band = atomicAdd(p.ticket, 1);
    if (threadIdx.x <= kMaxWarps) { s_avail[threadIdx.x] = 0; s_consumed[threadIdx.x] = 0; }
    __syncthreads();
    const int band = s_band;
    const long long band_r0 = 1 + (long long)band * wpc * 32;
    const int* prog_in = p.progress + band;          // written by band-1 ([0] is pre-armed)
    int* prog_out      = p.progress + band + 1;

    // CTA start gate: nobody spins on shared memory until the band above has produced
    // the first blocks of its last row.
    if (w == 0) {
        const int phi_prod = (int)(((band_r0 - 1) * p.pitch) & 3);
        const int qbp = (int)((p.m + phi_prod) >> 2) + 1;
        const int want = min(32, qbp);
        while (ld_acquire_gpu(prog_in) < want) __nanosleep(256);
    }
    __syncthreads();

    const long long r0 = band_r0 + 32LL * w;
    if (r0 > p.n) return;

    Strip<MU> S;
    S.lane = lane;
    const int phi0 = (int)((r0 * p.pitch) & 3);
    const int lm   = phi0 + lane * MU;
    const int sigma = lm >> 2;
    const int phil  = lm & 3;
    S.sigma31 = (phi0 + 31 * MU) >> 2;
    S.wrap0   = (phi0 < MU) ? 1 : 0;
    S.cbase   = -lane * (4 + MU) - phi0;
    S.m       = (int)p.m;
    const long long row = r0 + lane;
    const bool row_ok = row <= p.n;
    S.b4 = row_ok ? (unsigned)p.b[row - 1] * 0x01010101u : 0u;
    S.aw = p.a4 + (size_t)(3 - phil) * p.a4_stride + kAOff - lane - sigma;
    S.sm = p.s_match; S.sx = p.s_mismatch; S.gu = p.g_up; S.gl = p.g_left;
    S.stage    = smem4 + w * kWarpSmemBlocks;
    S.ring_in  = S.stage + 32 * kGroup;
    S.ring_out = S.ring_in + kWarpSmemBlocks;
    S.has_consumer = (w + 1 < wpc) && (r0 + 32 <= p.n);
    const bool band_last  = (w + 1 == wpc) && (r0 + 32 <= p.n);
    const bool src_global = (w == 0);
    S.fk = lane >> 3; S.fe = lane & 7;
    S.rowmask = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) S.rowmask |= (r0 + c + 8 * S.fk <= p.n) ? (1 << c) : 0;
    const long long Z = r0 * p.pitch - phi0;                 // multiple of 4
    S.q4 = (p.pitch - 4 - MU) >> 2;
    S.g0 = (Z >> 2) - 7 + S.fe + (long long)(8 * S.fk) * S.q4;
    S.fcol0 = 4 * (S.fe - 7) - (8 * S.fk) * (4 + MU) - phi0;
    S.H4 = reinterpret_cast<int4*>(p.H); S.P4 = reinterpret_cast<int4*>(p.P);
    S.H = p.H; S.P = p.P;
#pragma unroll
    for (int e = 0; e < 4; ++e) { S.A[e] = 0; S.B[e] = 0; }
    S.hl = 0; S.rmax = 0;
    S.aword = __ldg(S.aw);

    // fast (unmasked) steps: every lane's block inside columns [1, m] for steps t-7..t
    const int t_lo = (1 + 31 * (4 + MU) + phi0 + 3) >> 2;
    const int t_hi = (int)((p.m - 3 + phi0) >> 2);            // floor; may be negative

    // producer row (row r0-1) as seen by lane 0 when it comes from global memory
    const int phi_prod = (phi0 - MU) & 3;
    const int qbp = (int)((p.m + phi_prod) >> 2) + 1;
    const int4* Hprev4 = reinterpret_cast<const int4*>(p.H) + (((r0 - 1) * p.pitch) >> 2);
    int gl_loaded = 0;
    int cached_avail = 0;
    volatile int* v_avail = s_avail;
    volatile int* v_consumed = s_consumed;

    for (int tg = 0; tg < p.steps; tg += kGroup) {
        // ---- the blocks of the row above that this group will read: index <= tg+8+wrap0
        if (src_global) {
            const int need = min(tg + 9 + S.wrap0, qbp);
            while (gl_loaded < need) {
                const int want = min(gl_loaded + 32, qbp);
                while (ld_acquire_gpu(prog_in) < want) __nanosleep(64);
                const int blk = gl_loaded + lane;
                int4 v = make_int4(0, 0, 0, 0);
                if (blk < qbp) v = ld_cg_int4(Hprev4 + blk);
                S.ring_in[blk & (kRingBlocks - 1)] = make_int4(v.x << 4, v.y << 4, v.z << 4, v.w << 4);
                gl_loaded += 32;
                __syncwarp();
            }
        } else {
            const int need = min(tg + 9 + S.wrap0, p.qbmax);
            if (cached_avail < need) {
                do { cached_avail = v_avail[w]; } while (cached_avail < need);
                __threadfence_block();
            }
            if (lane == 0) v_consumed[w] = tg + 1 + S.wrap0;
        }
        if (tg == 0 && lane == 0) {
            const int4 va = S.ring_in[S.wrap0];
            S.A[0] = va.x; S.A[1] = va.y; S.A[2] = va.z; S.A[3] = va.w;
            if (S.wrap0) { const int4 vb = S.ring_in[0]; S.B[0] = vb.x; S.B[1] = vb.y; S.B[2] = vb.z; S.B[3] = vb.w; }
        }
        // ---- room in the next strip's ring for the 8 blocks lane 31 is about to write
        if (S.has_consumer) {
            const int last = tg + 7 - 31 - S.sigma31;
            if (last >= kRingBlocks)
                while (last - kRingBlocks >= v_consumed[w + 1]) { }
        }
        // ---- 8 steps
        if (tg - 7 >= t_lo && tg + 7 <= t_hi) {
#pragma unroll
            for (int j = 0; j < kGroup; ++j) S.template step<false>(tg + j);
        } else {
#pragma unroll 1
            for (int j = 0; j < kGroup; ++j) S.template step<true>(tg + j);
        }
        // ---- publish
        const int done = tg + 8 - 31 - S.sigma31;       // blocks of my last row finished so far
        if (S.has_consumer && done > 0 && lane == 31) {
            __threadfence_block();
            v_avail[w + 1] = done;
        }
        if (band_last && lane == 0) {
            // row 31 was written out at the step t == 6 (mod 8) of this group: blocks <= tg+6-31-sigma31
            const int vis = tg + 7 - 31 - S.sigma31;
            if (vis > 0) st_release_gpu(prog_out, vis);   // cumulative over the __syncwarp'd stores of the other lanes
        }
    }

    // ---- per-row maxima (omp_smithW.c:384-387 needs only the arg-max; see argmax_kernel)
    int hmax = row_ok ? (S.rmax >> 4) : 0;
    if (row_ok) p.row_max[row] = hmax;
    const int wm = __reduce_max_sync(0xffffffffu, hmax);
    if (lane == 0 && wm > 0) atomicMax(p.gmax, wm);
}

// ---------------------------------------------------------------------------------
// maxPos with the reference's tie-break: among the cells with H == global max, the one
// with the smallest i+j, then the largest i (omp_smithW.c:203-215,282-291,384-387).
// Only rows whose row maximum equals the global maximum are scanned.
// ---------------------------------------------------------------------------------
__global__ void argmax_kernel(const int32_t* __restrict__ H, long long pitch, long long m, long long n,
                              const int* __restrict__ row_max, const int* __restrict__ gmax,
                              unsigned long long* key)
{
    const int g = *gmax;
    if (g <= 0) return;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = 1 + warp; r <= n; r += nwarps) {
        if (row_max[r] != g) continue;
        const int32_t* Hr = H + r * pitch;
        for (long long j0 = 1; j0 <= m; j0 += 32) {
            const long long j = j0 + lane;
            const int v = (j <= m) ? Hr[j] : -1;
            const unsigned hit = __ballot_sync(0xffffffffu, v == g);
            if (hit) {
                if (lane == 0) {
                    const long long jj = j0 + (__ffs(hit) - 1);
                    const unsigned long long k = ((unsigned long long)(r + jj) << 32) |
                                                 (unsigned long long)(0xffffffffu - (unsigned)r);
                    atomicMin(key, k);
                }
                break;      // later columns of this row lie on later anti-diagonals
            }
        }
    }
}

__global__ void finalize_kernel(const unsigned long long* key, const int* gmax, long long pitch,
                                long long* maxPos, int32_t* maxScore)
{
    const int g = *gmax;
    long long pos = 0;
    if (g > 0) {
        const unsigned long long k = *key;
        const long long r = (long long)(0xffffffffu - (unsigned)(k & 0xffffffffu));
        const long long j = (long long)(k >> 32) - r;
        pos = r * pitch + j;
    }
    if (maxPos) *maxPos = pos;
    if (maxScore) *maxScore = g;
}

// ---------------------------------------------------------------------------------
// backtrack (omp_smithW.c:405-420): follow P from maxPos until a NONE cell, negating
// the path in place.  One warp: the 32x32 window of P ending at the current cell is
// fetched with 32 coalesced loads in flight, lane 0 walks inside it (>= 32 moves per
// window), so the serial chain pays one global round trip per window, not per cell.
// ---------------------------------------------------------------------------------
__global__ void backtrack_kernel(int32_t* P, long long pitch, long long maxPos_arg,
                                 const long long* d_maxPos, long long* d_pathLen)
{
    __shared__ int win[32][33];
    const int lane = threadIdx.x;
    long long pos = d_maxPos ? *d_maxPos : maxPos_arg;
    long long len = 0;
    if (pos > 0) {
        long long i = pos / pitch, j = pos % pitch;
        while (true) {
            const long long wi0 = i - 31, wj0 = j - 31;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
                const long long gi = wi0 + rr, gj = wj0 + lane;
                win[rr][lane] = (gi >= 0 && gj >= 0) ? P[gi * pitch + gj] : 0;
            }
            __syncwarp();
            int done = 0, li = 31, lj = 31;
            if (lane == 0) {
                while (li >= 0 && lj >= 0) {
                    const int pv = win[li][lj];
                    if (pv == 0) { done = 1; break; }                 // NONE ends the path (:419)
                    P[(wi0 + li) * pitch + (wj0 + lj)] = -pv;        // *= PATH (:417)
                    ++len;
                    if (pv == 3)      { --li; --lj; }                // DIAGONAL (:410)
                    else if (pv == 1) { --li; }                      // UP (:412)
                    else              { --lj; }                      // LEFT (:414)
                }
            }
            done = __shfl_sync(0xffffffffu, done, 0);
            li = __shfl_sync(0xffffffffu, li, 0);
            lj = __shfl_sync(0xffffffffu, lj, 0);
            i = wi0 + li; j = wj0 + lj;
            __syncwarp();
            if (done || i < 0 || j < 0) break;
        }
    }
    if (lane == 0 && d_pathLen) *d_pathLen = len;
}

}  // namespace swb
