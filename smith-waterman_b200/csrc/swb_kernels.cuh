// swb_kernels.cuh -- sm_100a kernels of the Smith-Waterman fill / backtrack path.
//
// Replaces the nDiag wavefront loop + similarityScore + backtrack of the reference
// (omp_smithW.c:203-216, 331-388, 405-420).  DESIGN.md has the derivation and the
// measurements; the short version of the fill kernel:
//
//  * The matrix is cut into STRIPS of 32 rows.  A COMPUTE warp owns a strip and sweeps it
//    left to right: lane l owns row r0+l and computes, per STEP t, the BLOCK j = t-l of four
//    columns 4j..4j+3 (column 0 is the zero boundary column and is computed like any other).
//    H stays in registers on the dependency chain: the block of the row above arrives by
//    __shfl_up_sync, the cell to the left is the lane's own previous value.
//  * A cell is computed on packed keys K = 16*H + tie, tie in {NONE 8, DIAG 7, UP 5, LEFT 2}:
//    one max over the four candidates reproduces the reference's strict-'>' order
//    DIAGONAL, UP, LEFT (omp_smithW.c:348-378); P = K&3, H = K>>4.  Three DPX VIADDMNMX per
//    cell, of which one is on the chain.
//  * Warp specialisation.  The compute warp only stages its packed block (one 16-byte
//    STS per lane per step) in a shared-memory ring.  A WRITER warp per strip drains the
//    ring: for every row it reads 32 consecutive columns (conflict-free LDS.32), unpacks H
//    and P and stores them with two warp-wide stores that each cover ONE FULL 128-byte
//    line of the caller's row-major matrix (the segmentation is chosen per row so that
//    this holds for any pitch).  The writer also keeps the strip maximum for maxPos.
//  * Strip -> strip hand-off (row 32 of a strip feeds row 1 of the next): lane 31 stores
//    its H block into a 64-entry shared-memory ring of the next compute warp of the CTA
//    ("band" = wpc strips); the entry carries an epoch tag in its low bits, so the
//    consumer polls the DATA and no fence sits on the dependency chain.  Band -> band goes
//    through a tagged boundary row in global memory (L2) that a LOADER warp of the next
//    band copies into that band's first ring.  Bands are claimed from an atomic ticket in
//    start order, so a waiting band's predecessor is always resident: the whole fill is
//    ONE launch, anti-diagonals are not launches.
//  * maxPos: a second tiny kernel scans only the strips that attain the global maximum,
//    with the reference's tie-break (first in anti-diagonal order, bottom-left to
//    top-right; omp_smithW.c:203-215,384-387).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace swb {

constexpr int kT        = 64;          // staging ring depth in steps (16-byte slots per row)
constexpr int kRowInts  = 4 * kT;      // ints per row of the staging ring
constexpr int kRing     = 64;          // hand-off ring capacity in blocks (power of two)
constexpr int kGroup    = 8;           // steps per synchronisation group
constexpr int kAPad     = 64;          // leading pad words of the packed copy of a
constexpr int kMaxWpc   = 4;           // strips per band (CTA) upper bound
constexpr int kDrainRounds = 2;        // writer rounds after the last compute group
// writer round r reads steps [8r-8, 8r+7]; compute group g overwrites the slots of group
// g - kT/8, which rounds <= g - kT/8 + 1 read: g may start once that many rounds are done
constexpr int kStageSlack = kT / kGroup - 2;

// tie codes: larger wins on equal score => NONE > DIAGONAL > UP > LEFT, and code&3 is
// the reference's P value (omp_smithW.c:33-36)
constexpr int kTieNone = 8, kTieDiag = 7, kTieUp = 5, kTieLeft = 2;

struct FillParams {
    const unsigned* a4;      // a4[kAPad + j] = a[4j-1 .. 4j+2] (block j; 0 outside [0,m))
    const unsigned char* b;  // n bytes, device
    int32_t*        H;
    int32_t*        P;
    long long       pitch;   // ints per row (>= m+1)
    long long       m, n;
    int             s_match, s_mismatch;   // 16*score + kTieDiag
    int             g_up, g_left;          // 16*gap + kTieUp / kTieLeft
    int             ngroups;               // compute groups per strip
    int             jmax;                  // last block holding a valid column: m >> 2
    int             wpc;                   // compute warps (strips) per band
    int4*           boundary;              // [nbands][bstride] tagged blocks: last row of band k
    long long       bstride;
    int*            ticket;                // band ticket counter
    int*            strip_max;             // [nstrips] max H of each strip
    int*            gmax;                  // global max H
    unsigned long long* trace;             // optional [nstrips][8] globaltimer stamps (developer tool) or nullptr
};

__device__ __forceinline__ void trace_stamp(const FillParams& p, long long strip, int slot, int lane)
{
    if (p.trace && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[strip * 8 + slot] = t;
    }
}

__device__ __forceinline__ int4 ld_cg_int4(const int4* p)
{
    int4 v;
    asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_cg_int4(int4* p, const int4& v)
{
    asm volatile("st.global.cg.v4.s32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cs_int(int32_t* p, int v)
{
    asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int4 lds_volatile_int4(const int4* p)
{
    int4 v;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_int4(int4* p, const int4& v)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.volatile.shared.v4.s32 [%0], {%1,%2,%3,%4};"
                 ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// stores under a PTX predicate: no branch, so a compute warp cannot diverge here
__device__ __forceinline__ void sts_volatile_int4_if(int4* p, const int4& v, bool on)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q st.volatile.shared.v4.s32 [%0], {%1,%2,%3,%4}; }"
                 ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((unsigned)on) : "memory");
}
__device__ __forceinline__ void st_cg_int4_if(int4* p, const int4& v, bool on)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q st.global.cg.v4.s32 [%0], {%1,%2,%3,%4}; }"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((unsigned)on) : "memory");
}

// ---------------------------------------------------------------------------------
// prep: packed copy of a -- word kAPad+j holds the characters of block j (columns
// 4j..4j+3; column c reads a[c-1], omp_smithW.c:395), zero outside the sequence --
// and arms the workspace words.
// ---------------------------------------------------------------------------------
__global__ void prep_kernel(const unsigned char* __restrict__ a, long long m,
                            unsigned* __restrict__ a4, long long nwords,
                            int* ticket, int* gmax, unsigned long long* key)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    for (long long w = tid; w < nwords; w += nth) {
        const long long base = 4 * (w - kAPad) - 1;
        unsigned word = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long idx = base + e;
            const unsigned c = (idx >= 0 && idx < m) ? (unsigned)a[idx] : 0u;
            word |= c << (8 * e);
        }
        a4[w] = word;
    }
    if (tid == 0) { *ticket = 0; *gmax = 0; *key = ~0ull; }
}

// ---------------------------------------------------------------------------------
// compute warp
// ---------------------------------------------------------------------------------
#ifndef SWB_X_GATESLEEP
#define SWB_X_GATESLEEP 100
#endif

// ---------------------------------------------------------------------------------
// compute warp.  IMPORTANT: a compute warp never diverges -- every poll loop is executed
// by all 32 lanes on a broadcast address.  (A lane-0-only spin loop leaves the warp split
// and every following shuffle takes the divergent slow path: measured 8x slower steps.)
// ---------------------------------------------------------------------------------
struct Strip {
    int lane;
    unsigned b4;
    int sm, sx, gu, gl;
    // dependency state (registers): block of the row above for this step, its last
    // element of the previous step (diagonal of my first column), my last cell
    int A0, A1, A2, A3, dgp, hl;
    int s0, s1, s2, s3;           // substitution scores of this step's four cells
    unsigned sa;                  // shared address of my next staging slot
    unsigned sa_base;             // my row's 1 KB staging region
    int4* ring_in;                // blocks of the row above my strip (tagged)
    int4* ring_out;               // blocks of my last row, for the next strip of the band
    int4* gout;                   // same, for the next band (global): slot of block t - lane
    bool  has_in;                 // a strip above exists
    bool  out_ring, out_glob;     // THIS LANE hands blocks on (lane 31 only): to the ring / to global
    int   jmax;

    __device__ __forceinline__ void scores(const unsigned aword)
    {
        const unsigned x = aword ^ b4;
        s0 = (x & 0x000000ffu) ? sx : sm;       // omp_smithW.c:394-399
        s1 = (x & 0x0000ff00u) ? sx : sm;
        s2 = (x & 0x00ff0000u) ? sx : sm;
        s3 = (x & 0xff000000u) ? sx : sm;
    }

    // one step.  EDGE: handles the zero column / not-yet-started lanes (first 4 groups) and
    // the end of the producer's row (last groups).  next_word: packed characters of the
    // NEXT step (its scores are computed in the shadow of the shuffles).
    template <bool EDGE>
    __device__ __forceinline__ void step(const int t, const unsigned next_word,
                                         const int4* in_next, const int want_next, const int out_slot,
                                         const int out_tag)
    {
        const int j = t - lane;
        const bool poll = has_in && (!EDGE || t + 1 <= jmax);
        // block t+1 of the strip above (lane 0's row above for the next step): first try early
        int4 v = make_int4(0, 0, 0, 0);
        if (poll) v = lds_volatile_int4(in_next);

        // K = max(left+gap|LEFT, up+gap|UP, diag+s|DIAG, 0|NONE)     (omp_smithW.c:339-381)
        const int p0 = __viaddmax_s32(dgp, s0, kTieNone);
        const int p1 = __viaddmax_s32(A0, s1, kTieNone);
        const int p2 = __viaddmax_s32(A1, s2, kTieNone);
        const int p3 = __viaddmax_s32(A2, s3, kTieNone);
        const int t0 = __viaddmax_s32(A0, gu, p0);
        const int t1 = __viaddmax_s32(A1, gu, p1);
        const int t2 = __viaddmax_s32(A2, gu, p2);
        const int t3 = __viaddmax_s32(A3, gu, p3);
        dgp = A3;
        int k0 = __viaddmax_s32(hl, gl, t0);
        if (EDGE) { if (j <= 0) k0 = kTieNone; }
        const int h0 = k0 & ~15;
        const int n0 = __shfl_up_sync(0xffffffffu, h0, 1);
        int k1 = __viaddmax_s32(h0, gl, t1);
        if (EDGE) { if (j < 0) k1 = kTieNone; }
        const int h1 = k1 & ~15;
        const int n1 = __shfl_up_sync(0xffffffffu, h1, 1);
        int k2 = __viaddmax_s32(h1, gl, t2);
        if (EDGE) { if (j < 0) k2 = kTieNone; }
        const int h2 = k2 & ~15;
        const int n2 = __shfl_up_sync(0xffffffffu, h2, 1);
        int k3 = __viaddmax_s32(h2, gl, t3);
        if (EDGE) { if (j < 0) k3 = kTieNone; }
        const int h3 = k3 & ~15;
        const int n3 = __shfl_up_sync(0xffffffffu, h3, 1);
        hl = h3;

        // ---------------- stage my packed block for the writer ----------------
#ifndef SWB_X_NOSTAGE
        asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(sa), "r"(k0), "r"(k1), "r"(k2), "r"(k3) : "memory");
#endif
        sa = ((sa + 16u) & (unsigned)(kT * 16 - 1)) | sa_base;

        // ---------------- hand my last row to the next strip (lane 31 only) ----------------
        {
            const bool started = !EDGE || j >= 0;
            sts_volatile_int4_if(ring_out + out_slot, make_int4(h0 | out_tag, h1, h2, h3 | out_tag), out_ring && started);
            st_cg_int4_if(gout, make_int4(h0 | 1, h1, h2, h3 | 1), out_glob && started);
            gout += 1;                                            // block j+1 next step
        }
        // ---------------- scores of the next step ----------------
        scores(next_word);

        // ---------------- the row above, for the next step ----------------
        if (poll) {
            int spins = 0;
            while (((v.x & 3) != want_next) | ((v.w & 3) != want_next)) {
                if (++spins > 32) __nanosleep(32);
                v = lds_volatile_int4(in_next);
            }
        }
        const bool l0 = (lane == 0);
        A0 = l0 ? (v.x & ~15) : n0;
        A1 = l0 ? v.y : n1;
        A2 = l0 ? v.z : n2;
        A3 = l0 ? (v.w & ~15) : n3;
    }
};

__device__ __forceinline__ void compute_strip(const FillParams& p, Strip& S, const unsigned* aw,
                                              volatile int* staged, volatile int* drained,
                                              volatile int* consumed_in, volatile int* consumed_out,
                                              const bool ring_consumer, const long long strip)
{
    const int lane = S.lane;
    trace_stamp(p, strip, 0, lane);
    unsigned cur[kGroup + 1], nxt[kGroup];
#pragma unroll
    for (int i = 0; i < kGroup; ++i) cur[i] = __ldg(aw + i);

    // first input block (block 0 -> ring index 32, epoch 0 -> tag 1); all lanes poll
    if (S.has_in) {
        int4 v = lds_volatile_int4(S.ring_in + 32);
        while (((v.x & 3) != 1) | ((v.w & 3) != 1)) { __nanosleep(SWB_X_GATESLEEP); v = lds_volatile_int4(S.ring_in + 32); }
        if (lane == 0) { S.A0 = v.x & ~15; S.A1 = v.y; S.A2 = v.z; S.A3 = v.w & ~15; }
    }
    trace_stamp(p, strip, 1, lane);
    S.scores(cur[0]);

    const int gtail = (p.jmax - 8) >> 3;          // groups g <= gtail: t+1 <= jmax for all their steps
    for (int g = 0; g < p.ngroups; ++g) {
        if (g == 4) trace_stamp(p, strip, 2, lane);
        if (g == 8) trace_stamp(p, strip, 3, lane);
        const int t0 = g * kGroup;
        // ---- staging ring space: the writer must have drained the slots this group overwrites
        if (g > kStageSlack) {
            while (*drained < g - kStageSlack) { }
        }
        // ---- hand-off ring space (blocks up to t0+7-31 are written in this group)
        if (ring_consumer && t0 - 80 > 0) { while (*consumed_out < t0 - 80) { } }
        if (S.has_in && lane == 0) *consumed_in = t0;
        // ---- sequence words of the next group
#pragma unroll
        for (int i = 0; i < kGroup; ++i) nxt[i] = __ldg(aw + t0 + kGroup + i);
        cur[kGroup] = nxt[0];

        // consumer side: block t -> ring index (t+32)&63, epoch ((t+32)>>6)&1
        const int4* in_base  = S.ring_in + ((t0 + 32) & (kRing - 1));
        const int4* in_wrap  = S.ring_in + ((t0 + 40) & (kRing - 1));
        const int   want     = 1 + (((t0 + 32) >> 6) & 1);
        const int   want_w   = 1 + (((t0 + 40) >> 6) & 1);
        // producer side: block t-31 -> ring index (t+1)&63, epoch ((t+1)>>6)&1
        const int   ob       = (t0 & (kRing - 1)) + 1;
        const int   ob_w     = (t0 + 8) & (kRing - 1);
        const int   otag     = 1 + ((t0 >> 6) & 1);
        const int   otag_w   = 1 + (((t0 + 8) >> 6) & 1);

        if (g >= 4 && g <= gtail) {
#pragma unroll
            for (int i = 0; i < kGroup; ++i) {
                const bool last = (i == kGroup - 1);
                S.template step<false>(t0 + i, cur[i + 1], last ? in_wrap : in_base + i + 1, last ? want_w : want,
                                       last ? ob_w : ob + i, last ? otag_w : otag);
            }
        } else {
#pragma unroll
            for (int i = 0; i < kGroup; ++i) {
                const bool last = (i == kGroup - 1);
                S.template step<true>(t0 + i, cur[i + 1], last ? in_wrap : in_base + i + 1, last ? want_w : want,
                                      last ? ob_w : ob + i, last ? otag_w : otag);
            }
        }
#pragma unroll
        for (int i = 0; i < kGroup; ++i) cur[i] = nxt[i];
        // ---- publish the staged group to the writer
        __syncwarp();
#ifndef SWB_X_NOFENCE
        __threadfence_block();
#endif
        if (lane == 0) *staged = g + 1;
    }
    trace_stamp(p, strip, 4, lane);
}

// ---------------------------------------------------------------------------------
// writer warp: drains the staging ring of one strip into H and P
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void writer_strip(const FillParams& p, const long long r0, const int lane,
                                             const int* stage /* [32][kRowInts] */, int4* rowtab, int* ftab,
                                             volatile int* staged, volatile int* drained, const long long strip)
{
    // per-row constants: round r flushes, for row l, the 32 columns 32r-E .. 32r-E+31 whose
    // first element sits on a 128-byte line of H (and P); E is the smallest such offset for
    // which lane l has finished those columns by the end of compute group r
    const long long row = r0 + lane;
    const int ph = (int)((row * p.pitch) & 31);
    const int d  = (lane + ((31 - ph) >> 2)) >> 3;
    const int E  = 32 * d + ph;
    const long long G0 = row * p.pitch - E;                      // multiple of 32
    const int F  = (8 * lane - E) & (kRowInts - 1);              // ring index of column c is (c + 8l) mod kRowInts
    const unsigned long long hb = (unsigned long long)(p.H + G0);
    const unsigned long long pb = (unsigned long long)(p.P + G0);
    rowtab[lane] = make_int4((int)(unsigned)hb, (int)(unsigned)(hb >> 32), (int)(unsigned)pb, (int)(unsigned)(pb >> 32));
    ftab[lane] = F;
    ftab[32 + lane] = E;
    const unsigned rowmask = __ballot_sync(0xffffffffu, row <= p.n);
    const int Emax = __reduce_max_sync(0xffffffffu, E);
    const int Emin = __reduce_min_sync(0xffffffffu, E);
    __syncwarp();

#ifndef SWB_X_FINETRACE
    trace_stamp(p, strip, 5, lane);
#endif
#ifdef SWB_X_NOWRITER
    return;
#endif
    int mx = 0;
    const int rounds = p.ngroups + kDrainRounds;
    const int m = (int)p.m;
    for (int r = 0; r < rounds; ++r) {
        const int need = min(r + 1, p.ngroups);
        if (*staged < need) {
            int spins = 0;
#ifndef SWB_X_WRITERSLEEP
#define SWB_X_WRITERSLEEP 64
#endif
            while (*staged < need) { if (++spins > 8) __nanosleep(SWB_X_WRITERSLEEP); }
        }
#ifndef SWB_X_NOFENCE
        __threadfence_block();
#endif
        const int v = 32 * r + lane;
        const bool interior = (rowmask == 0xffffffffu) && (32 * r - Emax >= 0) && (32 * r + 31 - Emin <= m);
        if (interior) {
#pragma unroll 8
            for (int l = 0; l < 32; ++l) {
                const int4 tb = rowtab[l];
                const int idx = (v + ftab[l]) & (kRowInts - 1);
                const int k = stage[l * kRowInts + idx];
                int32_t* hp = reinterpret_cast<int32_t*>(((unsigned long long)(unsigned)tb.y << 32) | (unsigned)tb.x) + v;
                int32_t* pp = reinterpret_cast<int32_t*>(((unsigned long long)(unsigned)tb.w << 32) | (unsigned)tb.z) + v;
                st_cs_int(hp, k >> 4);
                st_cs_int(pp, k & 3);
                mx = max(mx, k);
            }
        } else {
#pragma unroll 2
            for (int l = 0; l < 32; ++l) {
                const int4 tb = rowtab[l];
                const int idx = (v + ftab[l]) & (kRowInts - 1);
                const int c = v - ftab[32 + l];
                if (((rowmask >> l) & 1u) && c >= 0 && c <= m) {
                    const int k = stage[l * kRowInts + idx];
                    int32_t* hp = reinterpret_cast<int32_t*>(((unsigned long long)(unsigned)tb.y << 32) | (unsigned)tb.x) + v;
                    int32_t* pp = reinterpret_cast<int32_t*>(((unsigned long long)(unsigned)tb.w << 32) | (unsigned)tb.z) + v;
                    st_cs_int(hp, k >> 4);
                    st_cs_int(pp, k & 3);
                    mx = max(mx, k);
                }
            }
        }
        __syncwarp();
        if (lane == 0) *drained = r + 1;
    }
    // strip maximum (omp_smithW.c:384-387 needs only the arg-max; see argmax_kernel)
#ifndef SWB_X_FINETRACE
    trace_stamp(p, strip, 6, lane);
#endif
    const int hm = __reduce_max_sync(0xffffffffu, mx) >> 4;
    if (lane == 0) {
        p.strip_max[strip] = hm;
        if (hm > 0) atomicMax(p.gmax, hm);
    }
}

// ---------------------------------------------------------------------------------
// loader warp: copies the tagged last row of the band above from global memory (L2)
// into the first hand-off ring of this band
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void loader_band(const int4* src, const int nblocks, int4* ring, const int lane,
                                            volatile int* consumed)
{
    int base = 0;
    int idle = 0;
    while (base < nblocks) {
        const int limit = min(nblocks, *consumed + kRing);
        const int j = base + lane;
        bool ok = false;
        if (j < limit) {
            const int4 v = ld_cg_int4(src + j);
            ok = ((v.x & 3) == 1) && ((v.w & 3) == 1);
            if (ok) {
                const int tag = 1 + (((j + 32) >> 6) & 1);
                sts_volatile_int4(ring + ((j + 32) & (kRing - 1)),
                                  make_int4((v.x & ~15) | tag, v.y, v.z, (v.w & ~15) | tag));
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        const int lead = (mask == 0xffffffffu) ? 32 : (__ffs(~mask) - 1);
        base += lead;
        if (lead == 0) { if (++idle > 2) __nanosleep(idle > 64 ? 400 : 60); } else idle = 0;
    }
}

// ---------------------------------------------------------------------------------
// The fill kernel.  grid = number of bands, block = 32*(2*wpc+1) threads:
// warps [0,wpc) compute, [wpc,2wpc) write, warp 2wpc loads the band boundary.
// dynamic smem = wpc * (32*kRowInts*4 + kRing*16 + 32*16 + 64*4) bytes.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * (2 * kMaxWpc + 1))
fill_kernel(const FillParams p)
{
    extern __shared__ __align__(1024) int4 smem4[];
    __shared__ int s_band;
    __shared__ int s_staged[kMaxWpc], s_drained[kMaxWpc], s_consumed[kMaxWpc + 1];

    const int lane = threadIdx.x & 31;
    const int w    = threadIdx.x >> 5;
    const int wpc  = p.wpc;

    int4* stage4  = smem4;                                       // [wpc][32][kT]
    int4* rings   = stage4 + (size_t)wpc * 32 * kT;              // [wpc][kRing]
    int4* rowtabs = rings + (size_t)wpc * kRing;                 // [wpc][32]
    int*  ftabs   = reinterpret_cast<int*>(rowtabs + (size_t)wpc * 32);   // [wpc][64]

    if (threadIdx.x == 0) s_band = atomicAdd(p.ticket, 1);
    if (threadIdx.x < kMaxWpc) { s_staged[threadIdx.x] = 0; s_drained[threadIdx.x] = 0; }
    if (threadIdx.x <= kMaxWpc) s_consumed[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < wpc * kRing; i += blockDim.x) rings[i] = make_int4(0, 0, 0, 0);
    __syncthreads();
    const int band = s_band;
    const long long band_r0 = 1 + (long long)band * wpc * 32;

    if (w < wpc) {
        // ------------------------------------------------ compute
        const long long r0 = band_r0 + 32LL * w;
        if (r0 > p.n) return;
        Strip S;
        S.lane = lane;
        const long long row = r0 + lane;
        S.b4 = (row <= p.n) ? (unsigned)p.b[row - 1] * 0x01010101u : 0u;
        // keep the scoring constants in registers: a shuffle result is opaque to ptxas, which
        // otherwise re-reads them from the constant bank at the head of every step, on the
        // dependency chain
        S.sm = __shfl_sync(0xffffffffu, p.s_match, 0);
        S.sx = __shfl_sync(0xffffffffu, p.s_mismatch, 0);
        S.gu = __shfl_sync(0xffffffffu, p.g_up, 0);
        S.gl = __shfl_sync(0xffffffffu, p.g_left, 0);
        S.A0 = S.A1 = S.A2 = S.A3 = 0; S.dgp = 0; S.hl = 0;
        S.sa_base = (unsigned)__cvta_generic_to_shared(stage4 + ((size_t)w * 32 + lane) * kT);
        S.sa = S.sa_base + 16u * (unsigned)lane;                 // slot (t + lane) & (kT-1) at t = 0
        S.ring_in  = rings + (size_t)w * kRing;
        S.ring_out = rings + (size_t)(w + 1 < wpc ? w + 1 : w) * kRing;
        S.jmax = p.jmax;
        S.has_in = (r0 > 1);
        const bool next_row = (r0 + 32 <= p.n);                  // a strip below exists
        const bool ring_consumer = next_row && (w + 1 < wpc);
        S.out_ring = ring_consumer && lane == 31;
        S.out_glob = next_row && (w + 1 == wpc) && lane == 31;
        // block j = t - lane of step t goes to gout[j]; the pointer advances one block per step
        // (only lane 31's copy is ever dereferenced, from t = 31 on)
        S.gout = p.boundary + (size_t)(next_row && (w + 1 == wpc) ? band : 0) * p.bstride - lane;
        const unsigned* aw = p.a4 + kAPad - lane;                // aw[t] = characters of block t - lane
        compute_strip(p, S, aw, s_staged + w, s_drained + w, s_consumed + w, s_consumed + w + 1, ring_consumer,
                      (r0 - 1) >> 5);
    } else if (w < 2 * wpc) {
        // ------------------------------------------------ writer
        const int cw = w - wpc;
        const long long r0 = band_r0 + 32LL * cw;
        if (r0 > p.n) return;
        writer_strip(p, r0, lane, reinterpret_cast<const int*>(stage4 + (size_t)cw * 32 * kT),
                     rowtabs + (size_t)cw * 32, ftabs + (size_t)cw * 64, s_staged + cw, s_drained + cw,
                     (r0 - 1) >> 5);
    } else {
        // ------------------------------------------------ loader
        if (band == 0 || band_r0 > p.n) return;
        loader_band(p.boundary + (size_t)(band - 1) * p.bstride, p.jmax + 1, rings, lane, s_consumed);
    }
}

// ---------------------------------------------------------------------------------
// maxPos with the reference's tie-break: among the cells with H == global max, the one
// with the smallest i+j, then the largest i (omp_smithW.c:203-215,282-291,384-387).
// Only rows of strips whose maximum equals the global maximum are scanned.
// ---------------------------------------------------------------------------------
__global__ void argmax_kernel(const int32_t* __restrict__ H, long long pitch, long long m, long long n,
                              const int* __restrict__ strip_max, const int* __restrict__ gmax,
                              unsigned long long* key)
{
    const int g = *gmax;
    if (g <= 0) return;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = 1 + warp; r <= n; r += nwarps) {
        if (strip_max[(r - 1) >> 5] != g) continue;
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
