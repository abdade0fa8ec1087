// swb_kernels.cuh -- sm_100a kernels of the Smith-Waterman fill / backtrack path.
//
// Replaces the nDiag wavefront loop + similarityScore + backtrack of the reference
// (omp_smithW.c:203-216, 331-388, 405-420).  DESIGN.md has the derivation and the
// measurements; the short version of the fill kernel:
//
//  * The matrix is cut into STRIPS of 32*kR rows.  A COMPUTE warp owns a strip and sweeps it
//    left to right: lane l owns the kR adjacent rows r0+kR*l .. and computes, per STEP t, the
//    BLOCK j = t-l of four columns 4j..4j+3 of each of them (column 0 is the zero boundary
//    column and is computed like any other).  H stays in registers on the dependency chain:
//    the block of the row above the lane's first row arrives by __shfl_up_sync, the rows of
//    a lane feed each other directly, the cell to the left is the lane's own previous value.
//  * A cell is computed on packed keys K = 16*H + tie, tie in {NONE 8, DIAG 7, UP 5, LEFT 2}:
//    one max over the four candidates reproduces the reference's strict-'>' order
//    DIAGONAL, UP, LEFT (omp_smithW.c:348-378); P = K&3, H = K>>4.  Three DPX VIADDMNMX per
//    cell, of which one is on the chain.
//  * Warp specialisation.  The compute warp only stages its packed blocks (one 16-byte
//    STS per row per step) in a shared-memory ring.  WRITER warps drain the ring: for
//    every row they read 32 consecutive columns (conflict-free LDS.32), unpack H and P and
//    store them with two warp-wide stores that each cover ONE FULL 128-byte line of the
//    caller's row-major matrix (the segmentation is chosen per row so that this holds for
//    any pitch; a half-warp per row with 8-byte stores in the steady rounds).  The strip maximum
//    for maxPos is kept by the writers (single pair) or the compute warp (batch, score only).
//  * Strip -> strip hand-off (last row of a strip feeds the first row of the next): lane
//    31 stores its H block into a 64-entry shared-memory ring of the next compute warp of
//    the CTA ("band" = wpc strips); the entry carries an epoch tag in its low bits, so the
//    consumer polls the DATA and no fence sits on the dependency chain.  Band -> band goes
//    through a tagged boundary row in global memory (L2) that a LOADER warp of the next
//    band copies into that band's first ring.  Bands are claimed from an atomic ticket in
//    start order, so a waiting band's predecessor is always resident: the whole fill is
//    ONE launch, anti-diagonals are not launches.
//  * A compute warp never diverges: every poll is executed by all 32 lanes on a broadcast
//    address and hand-off stores are PTX-predicated.  (A lane-0-only spin loop leaves the
//    warp split and every following shuffle takes the divergent slow path: measured 8x
//    slower steps.)  It polls the strip above once per RUN of 4 or 8 steps, for the last block
//    the run needs, and executes the run without a branch (a tag check per step cost a third of
//    the step: the branch kept ptxas from overlapping consecutive steps); its waits never sleep
//    (__nanosleep oversleeps by microseconds and put whole launches into a slow mode).
//  * The edges of the matrix need no special code in the common case: the packed copy of a is
//    padded with the byte 0, which matches nothing unless b holds a NUL byte (the prep kernel
//    checks), so column 0, the blocks of lanes that have not started yet and the columns past m
//    evaluate to H = 0 / harmless values by themselves; a forced variant of the step covers the
//    NUL case, the column-strip mode (one pair on several GPUs) and the score-only tail.
//  * maxPos: a second tiny kernel scans only the strips that attain the global maximum,
//    with the reference's tie-break (first in anti-diagonal order, bottom-left to
//    top-right; omp_smithW.c:203-215,384-387).
// This header is included twice by swb_api.cu: as namespace swb (two rows per lane: batches, score-only) and as
// namespace swb_tall (three rows per lane: single large pairs) -- SWB_NS / SWB_ROWS_PER_LANE / SWB_FILL_ONLY select.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#ifndef SWB_X_GATESLEEP
#define SWB_X_GATESLEEP 0              // (no sleep at the gate either: see wait_block)
#endif
#ifndef SWB_X_LOADERSPINS
#define SWB_X_LOADERSPINS 512
#endif
#ifndef SWB_X_GT_LO
#define SWB_X_GT_LO 8                   // SWB_X_GROUPTRACE: groups [LO, HI) are accumulated as "steady"
#define SWB_X_GT_HI 0x7fffffff
#endif
#ifndef SWB_X_WRITERSLEEP
#define SWB_X_WRITERSLEEP 64
#endif

#ifndef SWB_STAGE_LATE
#define SWB_STAGE_LATE 0               // 1: the staging stores of a step are issued after its shuffles
#endif
#ifndef SWB_ST_MODE
#define SWB_ST_MODE 0
#endif
#ifdef SWB_ST
#undef SWB_ST
#endif
#if SWB_ST_MODE == 0
#define SWB_ST(p, v) __stcs(p, v)
#elif SWB_ST_MODE == 1
#define SWB_ST(p, v) __stcg(p, v)
#elif SWB_ST_MODE == 2
#define SWB_ST(p, v) (*(p) = (v))
#else
#define SWB_ST(p, v) __stwt(p, v)
#endif

#ifndef SWB_NS
#define SWB_NS swb
#endif
namespace SWB_NS {

#ifndef SWB_ROWS_PER_LANE
#define SWB_ROWS_PER_LANE 2
#endif
constexpr int kR        = SWB_ROWS_PER_LANE;   // adjacent rows per lane
constexpr int kStripRows = 32 * kR;    // rows per strip (compute warp)
#ifndef SWB_WRITER_ROWS
#define SWB_WRITER_ROWS 32
#endif
constexpr int kWRows    = SWB_WRITER_ROWS;         // rows of a strip drained by one writer warp (8, 16 or 32)
constexpr int kWriters  = kStripRows / kWRows;     // writer warps per strip
// (Tried and removed: writers that unpack into shared memory and store with 128-byte TMA bulk copies, one per row and
// matrix -- 12.3 ms against 5.1 ms for the 45000x45000 fill: bulk copies this small cost far more than the STGs.)
// KT (template parameter of the fill kernel) = staging ring depth in steps (16-byte slots per
// row): 64 for single large pairs, 32 for batches of small pairs (more CTAs per SM)
#ifndef SWB_RING_LOG
#define SWB_RING_LOG 6
#endif
constexpr int kRingLog  = SWB_RING_LOG;
constexpr int kRing     = 1 << kRingLog;   // hand-off ring capacity in blocks (power of two, >= 64)
constexpr int kGroup    = 8;           // steps per synchronisation group
#ifndef SWB_WAIT_STEPS
#define SWB_WAIT_STEPS 4
#endif
constexpr int kWaitSteps = SWB_WAIT_STEPS;   // steps per poll of the strip above (2, 4 or 8); the single-pair full fill uses 8
#ifndef SWB_WAIT_STEPS_SINGLE
#define SWB_WAIT_STEPS_SINGLE 8        // measured with the half skew (profiles/r02u_wait_steps.txt), 8 / 4 / 2 steps per poll:
#endif                                 // 45000x45000 4.29 / 4.32 / 4.48 ms, 100000x100000 17.8 / 17.9 / 18.3 ms, 2 000 000 columns
                                       // x 1000 rows 52.5 / 58.5 / 65.6 ms; only the chain-bound 1000 x 2 000 000 gains (75.5 / 69.3 /
                                       // 69.7 ms).  Score-only and batches: 4 beats 2 by 4-10 % and 8 by 2 %.
// Half skew (SWB_HALF_SKEW, the single-pair geometry): lane l trails lane l-1 by TWO columns instead of four.  A step
// still covers four columns of each of the lane's rows, but in two halves: the row above the first two columns was
// produced by the lane above in the second half of ITS previous step, the row above the last two in the first half of
// its CURRENT step (one pair of shuffles after each half).  Lane 31 then trails lane 0 by 16 steps instead of 31, so a
// strip can follow the strip above after 16 + 8 steps instead of 32 + 8: the strip-to-strip chain that bounds large
// fills shrinks by 40 % for the same instructions per step.  Odd lanes work on columns
// 4j+2 .. 4j+5: the packed copy of a exists twice, the second copy shifted by two columns.
#ifndef SWB_HALF_SKEW
#define SWB_HALF_SKEW 0
#endif
constexpr bool kHS        = SWB_HALF_SKEW != 0;
constexpr int kSkewCols   = kHS ? 2 : 4;           // columns a lane trails the lane above
constexpr int kSkewSteps  = kHS ? 16 : 31;         // steps lane 31 trails lane 0 (rounded up)
constexpr int kInOfs      = kHS ? -15 : 1;         // entry of the strip above that step t loads: t + kInOfs (entry x = the
                                                   // block lane 31 of the strip above produced in its step x + 31)
constexpr int kFirstEntry = kHS ? -16 : 0;         // first entry a strip needs (half skew: the row above columns 0 and 1)
constexpr int kAPad     = 64;          // leading pad words of the packed copy of a
// Score look-up (PROF instantiations): when b uses at most kMaxLetters distinct byte values and the scores fit a signed
// byte, the substitution scores are not derived from the characters with a compare and a select per cell.  The packed
// copy of a then holds, per block of four columns, one SELECTOR word: nibble e = the rank of column e's character among
// the letters of b (kPadCode for positions outside the sequence and for characters that do not occur in b).  Every lane
// keeps, for each of its rows, the eight score bytes of that row's character against the eight codes (two registers);
// ONE byte permute (PRMT) with the selector yields the row's four score bytes of a step, and the diagonal candidate of a
// cell is ONE dot-product instruction, kd = dp4a(scores, 16 << 8e, 16*H_diag + 7): IDP.4A issues on the FMA pipe
// (measured: tools/ubench_dp4a.cu), next to the IMAD that forms the UP candidate, so that the ALU pipe sees three
// instructions per cell (two three-input maxima and the mask) instead of 6.5.  H is carried as 16*H + kH7 in these
// instantiations: the tie code of the diagonal candidate then comes with the operand.
constexpr int kMaxLetters = 7;
constexpr int kPadCode    = 7;
constexpr int kBoundaryPad = 32;       // spare blocks in front of a band-boundary row (blocks -31..-1 of a strip's first steps)
constexpr int kMaxWpc   = 2;           // strips per band (CTA) upper bound: compute warps on schedulers 0..wpc-1,
                                       // writers + loader on the others
constexpr int kDrainRounds = 2;        // writer rounds after the last compute group
constexpr unsigned kGateSpinLimit = 40u << 20;   // polls (>= 200 ns each: ~10 s or more) a strip waits for its left GPU before it traps
#ifndef SWB_KMAX_IN_COMPUTE
#define SWB_KMAX_IN_COMPUTE 1          // who keeps the strip maximum: the compute warp (1) or its writers (0); the single-pair
                                       // full fill (KT == 64) always leaves it to the writers: 1 % faster there, while the
                                       // batch instantiation spills registers with it (SWB_KMAXC below)
#endif
#define SWB_KMAXC (SWB_KMAX_IN_COMPUTE && KT != 64)
// writer round r reads steps [8r-8, 8r+7]; compute group g overwrites the slots of group
// g - KT/8, which rounds <= g - KT/8 + 1 read: g may start once that many rounds are done
__host__ __device__ constexpr int stage_slack(int KT) { return KT / kGroup - 2; }

// tie codes: larger wins on equal score => NONE > DIAGONAL > UP > LEFT, and code&3 is
// the reference's P value (omp_smithW.c:33-36)
constexpr int kTieNone = 8, kTieDiag = 7, kTieUp = 5, kTieLeft = 2;
constexpr int kSNeg = -(1 << 30);      // forced substitution score: the diagonal candidate cannot win (16*H < 2^30)
constexpr int kHandOff = 5;            // P code of local column 0 in column-strip mode (not a reference code)

struct FillParams {
    const unsigned* a4;      // a4[kAPad + j] = a[4j-1 .. 4j+2] (block j; 0 outside [0,m))
    const unsigned* a4s;     // the same shifted by two columns: a4s[kAPad + j] = a[4j+1 .. 4j+4] (half skew: odd lanes)
    int             in_last; // last entry of the row above a strip that holds a column <= m
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
    int4*           boundary;              // [nbands][bstride] tagged blocks: last row of band k (kBoundaryPad spare blocks first)
    long long       bstride;
    int*            ticket;                // band ticket counter
    const int*      nul_flag;              // != 0: b holds a NUL byte (it would match the zero padding of a: see prep_kernel)
    // score look-up (see kMaxLetters)
    const unsigned char* lmap;             // [256] byte value -> its rank among the letters of b (all pairs), kPadCode if none
    const int*      nletters;              // distinct byte values in b; > kMaxLetters: the PROF instantiation returns at once,
                                           // <= kMaxLetters: the character-compare instantiation does
    int             match8, mismatch8;     // the raw scores (they fit a signed byte when prof_ok)
    int             prof_ok;               // host: the scores fit a byte (else the character-compare instantiation runs)
    int*            strip_max;             // [nstrips] max H of each strip (atomicMax by its writers)
    int*            gmax;                  // global max H
    unsigned long long* trace;             // optional [nstrips][8] globaltimer stamps (developer tool) or nullptr
    // batches of equally shaped pairs: pair k uses a4 + k*a4_stride, b + k*n, H/P + k*pair_stride,
    // boundary + k*(nbands-1)*bstride, strip_max + k*nstrips, gmax + k, row_best + k*(n+1)
    int             nbands;                // bands per pair
    long long       npairs;                // pairs in this launch
    long long       nstrips;               // strips per pair
    long long       a4_stride, pair_stride;
    // score-only mode (no H/P stores): per-row best cell, packed (score << 32) | (0xffffffff - column)
    unsigned long long* row_best;
    // column-strip mode (one GPU of several working on one pair): this GPU owns columns c0+1..c0+m of
    // the pair; local column 0 is the last column of the GPU to the left.
    const int32_t*  left_in;               // [n+1] H of that column (written by the left GPU), or nullptr
    const int*      left_flags;            // [ceil(n/kWRows)] == epoch once rows kWRows*k+1 .. kWRows*(k+1) are in left_in
    int32_t*        right_out;             // PEER pointer: the right GPU's left_in, or nullptr
    int*            right_flags;           // PEER pointer: the right GPU's left_flags
    int             epoch;
};

// Developer builds only (-DSWB_TRACE, `make trace` -> build/libswb200_trace.so, used by tools/trace.py and friends):
// the product library ignores swb_tuning.trace -- the checks alone were ~20 instructions per group of 8 steps on the
// compute warps.
#ifndef SWB_TRACE
#if defined(SWB_X_GROUPTRACE) || defined(SWB_X_WRITERTRACE) || defined(SWB_X_CLKTRACE)
#define SWB_TRACE 1
#else
#define SWB_TRACE 0
#endif
#endif
__device__ __forceinline__ void trace_stamp(const FillParams& p, long long strip, int slot, int lane)
{
    if (SWB_TRACE && p.trace && lane == 0) {
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
__device__ __forceinline__ void st_cs_int(int32_t* p, int v)
{
    asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// shared-memory accesses by 32-bit shared address + compile-time byte offset
template <int OFF>
__device__ __forceinline__ int4 lds_volatile_int4(unsigned a)
{
    int4 v;
    asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4+%5];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a), "n"(OFF) : "memory");
    return v;
}
__device__ __forceinline__ int4 lds_volatile_int4(const int4* p)
{
    return lds_volatile_int4<0>((unsigned)__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void sts_volatile_int4(int4* p, const int4& v)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.volatile.shared.v4.s32 [%0], {%1,%2,%3,%4};"
                 ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// stores under a PTX predicate: no branch, so a compute warp cannot diverge here
template <int OFF>
__device__ __forceinline__ void sts_volatile_int4_if(unsigned a, const int4& v, int on)
{
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %5, 0; @q st.volatile.shared.v4.s32 [%0+%6], {%1,%2,%3,%4}; }"
                 ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(on), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void st_cg_int4_if(const int4* p, const int4& v, int on)
{
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %5, 0; @q st.global.cg.v4.s32 [%0+%6], {%1,%2,%3,%4}; }"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(on), "n"(OFF) : "memory");
}
__device__ __forceinline__ void st_cg_int_if(int32_t* p, int v, int on)
{
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.global.cg.s32 [%0], %1; }" ::"l"(p), "r"(v), "r"(on) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts_int4(unsigned a, int x, int y, int z, int w)
{
    asm volatile("st.shared.v4.s32 [%0+%5], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w), "n"(OFF) : "memory");
}
__device__ __forceinline__ int lds_volatile_int(unsigned a)
{
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_int_if(unsigned a, int v, int on)
{
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.volatile.shared.s32 [%0], %1; }" ::"r"(a), "r"(v), "r"(on) : "memory");
}
// Wait until the shared-memory word at `a` is >= want.  The first check sits outside the
// loop: ptxas puts a YIELD into every loop with a volatile load, and a YIELD on the straight
// path costs ~100 clk even when the flag is already set.
__device__ __forceinline__ void spin_until_ge(unsigned a, int want)
{
#ifdef SWB_X_SPINFIRST
    if (__builtin_expect(lds_volatile_int(a) >= want, 1)) return;
#endif
    while (lds_volatile_int(a) < want) { }
}
// wait until block x of the strip above is in the hand-off ring (block x -> entry (x+32)&63, epoch tag 1 + ((x+32)>>6 & 1));
// all lanes poll the same address
__device__ __forceinline__ void wait_block(unsigned ring_in, int x)
{
    const unsigned a = ring_in + 16u * (unsigned)((x + 32) & (kRing - 1));
    const int want = 1 + (((x + 32) >> kRingLog) & 1);
    // Pure spin: __nanosleep oversleeps by microseconds on this hardware whatever its argument.  With a
    // `nanosleep(20)` after 64 polls here a strip that followed its producer closely fell into a mode of one sleep
    // per run of steps (1400 clk per step instead of 250) and dragged every later strip with it: score-only launches
    // took 12 or 18 ms instead of 5.8, at random.
    int v = lds_volatile_int(a);
    while ((v & 3) != want) v = lds_volatile_int(a);
}
__device__ __forceinline__ unsigned long long mad_wide(unsigned a, unsigned b, unsigned long long c)
{
    unsigned long long d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
// a value the compiler / ptxas cannot rematerialise from the constant bank
__device__ __forceinline__ int opaque(int x) { return __shfl_sync(0xffffffffu, x, 0); }

// ---------------------------------------------------------------------------------
// prep: packed copy of a -- word kAPad+j holds the characters of block j (columns
// 4j..4j+3; column c reads a[c-1], omp_smithW.c:395), zero outside the sequence --
// and arms the workspace words.
// ---------------------------------------------------------------------------------
__global__ void prep_kernel(const unsigned char* __restrict__ a, long long m, long long npairs,
                            unsigned* __restrict__ a4, long long nwords,
                            int* ticket, int* gmax, unsigned long long* key,
                            int* strip_max, long long nstrips_total,
                            const unsigned char* __restrict__ b, long long nb, int* nul_flag, int* present)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    // (two copies: the second one, nwords * npairs words further on, is shifted by two columns -- half skew)
    for (long long x = tid; x < 2 * nwords * npairs; x += nth) {
        const long long y = x % (nwords * npairs), shift = 2 * (x / (nwords * npairs));
        const long long pair = y / nwords, w = y % nwords;
        const unsigned char* ap = a + pair * m;
        const long long base = 4 * (w - kAPad) - 1 + shift;
        unsigned word = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long idx = base + e;
            const unsigned c = (idx >= 0 && idx < m) ? (unsigned)ap[idx] : 0u;
            word |= c << (8 * e);
        }
        a4[x] = word;
    }
    // The padding of a is the byte 0.  Unless b holds a NUL byte (never, for the reference's C strings) a padded
    // position matches nothing, and then -- with mismatch <= 0 and gap < 0 -- column 0, the blocks in front of it
    // (lanes that have not started yet) and the columns past m evaluate to H = 0 / values below their valid
    // neighbours by themselves: the compute warp runs the same code on every step of a strip.
    {
        bool nul = false;
        for (long long k = tid; k < nb; k += nth) {
            const unsigned c = b[k];
            nul |= (c == 0);
            if (present[c] == 0) present[c] = 1;       // alphabet of b (benign race: every writer stores 1)
        }
        if (nul) *nul_flag = 1;
    }
    for (long long k = tid; k < nstrips_total; k += nth) strip_max[k] = 0;
    for (long long k = tid; k < npairs; k += nth) { gmax[k] = 0; key[k] = ~0ull; }
    if (tid == 0) *ticket = 0;
}

// ---------------------------------------------------------------------------------
// score look-up (matchMissmatchScore, omp_smithW.c:394-399): ranks the byte values that occur in b (present[], set by
// prep_kernel; block 0 publishes the map and the count) and, with at most kMaxLetters of them, rewrites the packed copy
// of a as selector words -- nibble e of word kAPad+j = rank of a[4j+e-1], kPadCode outside the sequence or for a
// character that b does not use (it matches no row).
// ---------------------------------------------------------------------------------
__global__ void selector_kernel(const unsigned char* __restrict__ a, long long m, long long npairs,
                                const int* __restrict__ present, unsigned char* lmap, int* nletters,
                                unsigned* __restrict__ a4, long long nwords)
{
    __shared__ unsigned char s_map[256];
    __shared__ int s_cnt[8];
    __shared__ int s_n;
    {
        // rank of byte value c among the present ones (256 threads = one per value)
        const int c = threadIdx.x;
        const bool on = present[c] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if ((c & 31) == 0) s_cnt[c >> 5] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 8; ++w) { if (w < (c >> 5)) before += s_cnt[w]; total += s_cnt[w]; }
        const int rank = before + __popc(bal & ((1u << (c & 31)) - 1u));
        s_map[c] = (unsigned char)((on && rank < kMaxLetters) ? rank : kPadCode);
        if (c == 0) s_n = total;
        if (blockIdx.x == 0) lmap[c] = s_map[c];
        if (blockIdx.x == 0 && c == 0) *nletters = total;
        __syncthreads();
    }
    if (s_n > kMaxLetters) return;         // the characters stay in a4: the character-compare instantiation runs
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    for (long long x = tid; x < 2 * nwords * npairs; x += nth) {
        const long long y = x % (nwords * npairs), shift = 2 * (x / (nwords * npairs));
        const long long pair = y / nwords, w = y % nwords;
        const unsigned char* ap = a + pair * m;
        const long long base = 4 * (w - kAPad) - 1 + shift;
        unsigned word = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long idx = base + e;
            const unsigned code = (idx >= 0 && idx < m) ? (unsigned)s_map[ap[idx]] : (unsigned)kPadCode;
            word |= code << (4 * e);
        }
        a4[x] = word;
    }
}

// ---------------------------------------------------------------------------------
// compute warp
// ---------------------------------------------------------------------------------
template <int KT, bool STORE, bool PROF>
struct Strip {
    static constexpr int kH7 = PROF ? kTieDiag : 0;      // low bits of every H-carrying register (16*H + kH7)
    static constexpr int kRowInts = 4 * KT;
    int lane;
    unsigned b4[kR];              // my rows' characters, replicated in the four bytes
    unsigned inv[kR];             // 0, or 0x01010101 for a row past n (partial last strip): such a row never matches
    int sm, sx, gu, gl, g16;
    // dependency state (registers): block of the row above my first row for this step, its
    // last element of the previous step (diagonal of my first column), the last cell of
    // each of my rows
    int A0, A1, A2, A3, dgp, hl[kR];
    int s[kR][4];                 // substitution scores of this step's cells
    unsigned sa;                  // shared address of my next staging slot (first row; row q at +q KB)
    unsigned sa_base;             // my first row's 1 KB staging region
    unsigned ring_in;             // shared address: blocks of the row above my strip (tagged)
    unsigned ring_out;            // shared address: blocks of my last row, for the next strip of the band
    const int4* gout;             // same, for the next band (global); block j of this group's first step
    int   has_in;                 // a strip above exists
    int   out_ring, out_glob;     // THIS LANE hands blocks on (lane 31 only): to the ring / to global
    int   jmax;
    int   mcols;                  // m
    int   lb[kR];                 // 16*(H - gap) of my rows' local column 0 (column-strip mode: H of the left GPU's
                                  // last column; otherwise 16*(0 - gap)): injected as the left neighbour of block 0
    // score-only: best cell of each of my rows.  rbest = 16*H | (3 - column in its block) of the first column that reached
    // the row's maximum, rmax = rbest | 15 (what a later block must exceed: strictly larger H), rcol = the STEP it was found in
    int   rmax[kR], rcol[kR], rbest[kR];
    int   kmax;                   // full fill: largest key of my cells in columns 1..m (strip maximum; the writers have no
                                  // ALU slots to spare for it)
    // PROF: the score bytes of my rows' characters against codes 0..3 / 4..7, this step's four score bytes per row,
    // and an opaque 1 (multiplier of the IMADs that keep additions off the ALU pipe)
    unsigned tlo[kR], thi[kR];
    int   sb[kR];
    int   one;
    int   h7r;                    // kH7 in a register (LOP3 takes one immediate: (k & ~15) | 7 with two would be two instructions)

    // x + y on the FMA pipe (IMAD with a register multiplier that ptxas cannot fold)
    __device__ __forceinline__ int addf(const int x, const int y) const
    {
        int d;
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(y));
        return d;
    }
    // PROF: the score bytes of the step whose selector word is `sel` (one PRMT per row)
    __device__ __forceinline__ void lookup(const unsigned sel)
    {
#pragma unroll
        for (int q = 0; q < kR; ++q)          // (prmt.b32 directly: __byte_perm masks the selector with 0x7777 first)
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sb[q]) : "r"(tlo[q]), "r"(thi[q]), "r"(sel));
    }

    __device__ __forceinline__ void scores(const unsigned aword)
    {
#ifdef SWB_X_NOSCORES
        if (aword != 0x12345678u) return;          // timing experiment only: wrong results
#endif
#pragma unroll
        for (int q = 0; q < kR; ++q) {
            const unsigned x = (aword ^ b4[q]) | inv[q];
            s[q][0] = (x & 0x000000ffu) ? sx : sm;       // omp_smithW.c:394-399
            s[q][1] = (x & 0x0000ff00u) ? sx : sm;
            s[q][2] = (x & 0x00ff0000u) ? sx : sm;
            s[q][3] = (x & 0xff000000u) ? sx : sm;
        }
    }

    // Head of a strip (first 4 groups): a lane has not started while its block index jn is negative,
    // and column 0 (first cell of block 0) is the boundary column.  Nothing is forced on the dependency
    // chain: with all inputs zero and the substitution score forced to kSNeg a cell evaluates to
    // NONE / 0 by itself, and the boundary value of column 0 enters as the left neighbour.
    __device__ __forceinline__ void head_fix(const unsigned aword, const int jn)
    {
        // (overrides scores(): two selects per step instead of one per cell)
        const int sm0 = (jn <= 0) ? kSNeg : sm, sx0 = (jn <= 0) ? kSNeg : sx;
        const int sm1 = (jn < 0) ? kSNeg : sm, sx1 = (jn < 0) ? kSNeg : sx;
#pragma unroll
        for (int q = 0; q < kR; ++q) {
            const unsigned x = (aword ^ b4[q]) | inv[q];
            s[q][0] = (x & 0x000000ffu) ? sx0 : sm0;
            s[q][1] = (x & 0x0000ff00u) ? sx1 : sm1;
            s[q][2] = (x & 0x00ff0000u) ? sx1 : sm1;
            s[q][3] = (x & 0xff000000u) ? sx1 : sm1;
            hl[q] = (jn == 0) ? lb[q] : hl[q];
        }
    }

    // one step.  MODE 0: interior; bit 0: head of a strip (first 4 groups: the zero column, lanes that have not
    // started yet); bit 1: tail (last groups: the end of the producer's row, columns past m); 3: both (tiny matrices).  I = step index inside the group
    // (compile time: ring slots and the global hand-off are immediate offsets).
    // next_word: packed characters of the NEXT step (its scores are computed in the shadow
    // of the shuffles).  in_g / out_g: ring addresses of this group's first entries.
    template <int MODE, int I>
    __device__ __forceinline__ void step(const int t, const unsigned next_word,
                                         const unsigned in_g, const unsigned in_w, const int want, const int want_w,
                                         const unsigned out_g, const unsigned out_w, const int otag, const int otag_w)
    {
        constexpr bool LAST = (I == kGroup - 1);
        const int j = t - lane;
        // block t+1 of the strip above (the row above lane 0's first row in the next step): loaded early, it is needed
        // only after the shuffles.  Unconditional outside the tail: the first strip of a pair reads its zero-filled
        // ring (a predicated load cost eight extra moves per step), and past the end of the row above the ring holds
        // stale blocks that only reach columns > m.
        int4 v;
        if (MODE & 2) {
            v = make_int4(0, kH7, kH7, kH7);
            if (has_in && t + 1 <= jmax) v = LAST ? lds_volatile_int4<0>(in_w) : lds_volatile_int4<16 * (I + 1)>(in_g);
        } else {
            v = LAST ? lds_volatile_int4<0>(in_w) : lds_volatile_int4<16 * (I + 1)>(in_g);
        }

        // K = max(left+gap|LEFT, up+gap|UP, diag+s|DIAG, 0|NONE)     (omp_smithW.c:339-381)
        int u0 = A0, u1 = A1, u2 = A2, u3 = A3, dg = dgp;
        dgp = A3;
        int n0, n1, n2, n3;
        int kk[kR][4];
#pragma unroll
        for (int q = 0; q < kR; ++q) {
            int t0, t1, t2, t3;
            if constexpr (PROF) {
                // diagonal candidate: one IDP.4A per cell picks the cell's score byte, scales it by 16 and adds 16*H + 7;
                // UP candidate: one IMAD; both on the FMA pipe.  One three-input maximum folds them with the NONE key.
                int D0 = dg, D1 = u0, D2 = u1, D3 = u2;
                if (MODE & 1) {
                    // head of a strip in column-strip mode (see head_fix): blocks j < 0 and column 0 cannot take the
                    // diagonal, and the boundary value enters as the left neighbour of block 0
                    D0 = (j <= 0) ? kSNeg : D0; D1 = (j < 0) ? kSNeg : D1; D2 = (j < 0) ? kSNeg : D2; D3 = (j < 0) ? kSNeg : D3;
                    hl[q] = (j == 0) ? lb[q] : hl[q];
                }
                const int d0 = __dp4a(sb[q], 16, D0), d1 = __dp4a(sb[q], 16 << 8, D1);
                const int d2 = __dp4a(sb[q], 16 << 16, D2), d3 = __dp4a(sb[q], 16 << 24, D3);
                const int v0 = addf(u0, gu), v1 = addf(u1, gu), v2 = addf(u2, gu), v3 = addf(u3, gu);
                t0 = __vimax3_s32(d0, v0, kTieNone);
                t1 = __vimax3_s32(d1, v1, kTieNone);
                t2 = __vimax3_s32(d2, v2, kTieNone);
                t3 = __vimax3_s32(d3, v3, kTieNone);
            } else {
                const int p0 = __viaddmax_s32(dg, s[q][0], kTieNone);
                const int p1 = __viaddmax_s32(u0, s[q][1], kTieNone);
                const int p2 = __viaddmax_s32(u1, s[q][2], kTieNone);
                const int p3 = __viaddmax_s32(u2, s[q][3], kTieNone);
                t0 = __viaddmax_s32(u0, gu, p0);
                t1 = __viaddmax_s32(u1, gu, p1);
                t2 = __viaddmax_s32(u2, gu, p2);
                t3 = __viaddmax_s32(u3, gu, p3);
            }
            dg = hl[q];                                  // diagonal of the next row's first cell
            // (PROF: H travels as 16*H + 7, one LOP3 either way)
            const int k0 = __viaddmax_s32(hl[q], gl, t0);
            const int h0 = (k0 & ~15) | h7r;
            if (q == kR - 1) n0 = __shfl_up_sync(0xffffffffu, h0, 1);
            const int k1 = __viaddmax_s32(h0, gl, t1);
            const int h1 = (k1 & ~15) | h7r;
            if (q == kR - 1) n1 = __shfl_up_sync(0xffffffffu, h1, 1);
            const int k2 = __viaddmax_s32(h1, gl, t2);
            const int h2 = (k2 & ~15) | h7r;
            if (q == kR - 1) n2 = __shfl_up_sync(0xffffffffu, h2, 1);
            const int k3 = __viaddmax_s32(h2, gl, t3);
            const int h3 = (k3 & ~15) | h7r;
            if (q == kR - 1) n3 = __shfl_up_sync(0xffffffffu, h3, 1);
            hl[q] = h3;
            if (STORE && SWB_KMAXC) {
                int e0 = k0, e1 = k1, e2 = k2, e3 = k3;
                if (MODE & 1) e0 = (j == 0) ? 0 : k0;        // column 0 (blocks j < 0 hold NONE)
                if (MODE & 2) {
                    const int c = 4 * j;
                    e0 = (c <= mcols) ? e0 : 0;
                    e1 = (c + 1 <= mcols) ? k1 : 0;
                    e2 = (c + 2 <= mcols) ? k2 : 0;
                    e3 = (c + 3 <= mcols) ? k3 : 0;
                }
                kmax = __vimax3_s32(__vimax3_s32(kmax, e0, e1), e2, e3);
            }
            if (STORE) {
                // stage the packed block of this row for the writers (SWB_STAGE_LATE: after the last row's shuffles have
                // been issued, so that they do not queue behind the stores in the SM's load/store pipeline)
                if (SWB_STAGE_LATE) { kk[q][0] = k0; kk[q][1] = k1; kk[q][2] = k2; kk[q][3] = k3; }
                else {
                    if (q == 0) sts_int4<0>(sa, k0, k1, k2, k3);
                    if (q == 1) sts_int4<kRowInts * 4>(sa, k0, k1, k2, k3);
                    if (q == 2) sts_int4<2 * kRowInts * 4>(sa, k0, k1, k2, k3);
                    if (q == 3) sts_int4<3 * kRowInts * 4>(sa, k0, k1, k2, k3);
                }
            } else {
                // score only: remember the first column of this row that reaches its maximum
                // (strict '>' keeps the earliest column, i.e. the earliest anti-diagonal of the row);
                // branch-free -- the compute warp must not diverge
                // The column inside the block rides in the low four bits of the value (3 - e: among equal H the first
                // column is the larger key; four IMADs), one three-input maximum + one maximum find the block's best,
                // and it replaces the row's best only if its H is strictly larger (compare against best | 15).
                int v0 = addf(h0, 3 - kH7), v1 = addf(h1, 2 - kH7), v2 = addf(h2, 1 - kH7), v3 = addf(h3, 0 - kH7);
                if (MODE != 0) {
                    const int c = 4 * j;
                    v0 = (c >= 1 && c <= mcols) ? v0 : -1;
                    v1 = (c + 1 >= 1 && c + 1 <= mcols) ? v1 : -1;
                    v2 = (c + 2 >= 1 && c + 2 <= mcols) ? v2 : -1;
                    v3 = (c + 3 >= 1 && c + 3 <= mcols) ? v3 : -1;
                }
                const int bm = max(__vimax3_s32(v0, v1, v2), v3);
                const bool upd = bm > rmax[q];
                rbest[q] = upd ? bm : rbest[q];
                rmax[q] = upd ? (bm | 15) : rmax[q];
                rcol[q] = upd ? t : rcol[q];
            }
            u0 = h0; u1 = h1; u2 = h2; u3 = h3;          // the row above the next row
        }
        if (STORE && SWB_STAGE_LATE) {
#pragma unroll
            for (int q = 0; q < kR; ++q) {
                if (q == 0) sts_int4<0>(sa, kk[q][0], kk[q][1], kk[q][2], kk[q][3]);
                if (q == 1) sts_int4<kRowInts * 4>(sa, kk[q][0], kk[q][1], kk[q][2], kk[q][3]);
                if (q == 2) sts_int4<2 * kRowInts * 4>(sa, kk[q][0], kk[q][1], kk[q][2], kk[q][3]);
                if (q == 3) sts_int4<3 * kRowInts * 4>(sa, kk[q][0], kk[q][1], kk[q][2], kk[q][3]);
            }
        }
        if (STORE) sa = ((sa + 16u) & (unsigned)(KT * 16 - 1)) | sa_base;

        // ---------------- hand my last row to the next strip (lane 31 only) ----------------
        {
            const int started = 1;   // (blocks j < 0 go out as well: wrong-epoch ring entries / the spare blocks of the boundary row)
            (void)j;
            // (PROF: u0 ends in kH7 = 0111b and otag / otag_w arrive as tag ^ 7: one XOR leaves the tag in the low bits)
            const int4 o = make_int4(PROF ? (u0 ^ (LAST ? otag_w : otag)) : (u0 | (LAST ? otag_w : otag)), u1, u2, u3);
            if (LAST) sts_volatile_int4_if<0>(out_w, o, out_ring & started);
            else      sts_volatile_int4_if<16 * I>(out_g, o, out_ring & started);
            st_cg_int4_if<16 * I>(gout, o, out_glob & started);
        }
        // ---------------- scores of the next step ----------------
        if constexpr (PROF) lookup(next_word);
        else if (MODE & 1)  head_fix(next_word, j + 1);
        else                scores(next_word);

        // ---------------- the row above, for the next step ----------------
        // (v is valid by construction: compute_strip waited for the last block of this run of steps.  A tag
        // check with a spin loop HERE costs ~80 clk per step even when it never spins: the branch cuts the
        // group into basic blocks and ptxas can no longer overlap the tail of one step with the head of the next.)
        const bool l0 = (lane == 0);
        A0 = l0 ? (PROF ? ((v.x & ~15) | h7r) : (v.x & ~15)) : n0;
        A1 = l0 ? v.y : n1;
        A2 = l0 ? v.z : n2;
        A3 = l0 ? v.w : n3;
    }

    // The same step with HALF SKEW (kHS): my columns are c0 = 4t - 2*lane .. c0 + 3.  First half (c0, c0+1): the row
    // above is A0, A1 (shuffled at the end of the previous step); second half (c0+2, c0+3): the row above is what the
    // lane above produced in the first half of THIS step.  Lane 0 takes both from entry t + kInOfs of the strip above
    // (.xy now, .zw for the first half of the next step).  dgp = the row above at column c0 - 1.
    template <int MODE, int I>
    __device__ __forceinline__ void step_hs(const int t, const unsigned next_word,
                                            const unsigned in_g, const unsigned in_w, const int want, const int want_w,
                                            const unsigned out_g, const unsigned out_w, const int otag, const int otag_w)
    {
        constexpr bool LAST = (I == kGroup - 1);
        const int c0 = 4 * t - 2 * lane;
        int4 v;
        if (MODE & 2) {
            v = make_int4(0, kH7, kH7, kH7);
            if (has_in && t + kInOfs <= jmax) v = LAST ? lds_volatile_int4<0>(in_w) : lds_volatile_int4<16 * (I + 1)>(in_g);
        } else {
            v = LAST ? lds_volatile_int4<0>(in_w) : lds_volatile_int4<16 * (I + 1)>(in_g);
        }
        int hA0[kR], hA1[kR], kA0[kR], kA1[kR];
        int nA0, nA1, nB2, nB3;
        // ---------------- first half: columns c0, c0 + 1
        {
            int u0 = A0, u1 = A1, dg = dgp;
#pragma unroll
            for (int q = 0; q < kR; ++q) {
                int D0 = dg, D1 = u0, L0 = hl[q];
                if (MODE & 1) {
                    // head of a strip in column-strip mode: columns <= 0 cannot take the diagonal, and the boundary value
                    // enters as the left neighbour of column 0
                    D0 = (c0 <= 0) ? kSNeg : D0; D1 = (c0 + 1 <= 0) ? kSNeg : D1;
                    L0 = (c0 == 0) ? lb[q] : L0;
                }
                int t0, t1;
                if constexpr (PROF) {
                    const int d0 = __dp4a(sb[q], 16, D0), d1 = __dp4a(sb[q], 16 << 8, D1);
                    const int v0 = addf(u0, gu), v1 = addf(u1, gu);
                    t0 = __vimax3_s32(d0, v0, kTieNone);
                    t1 = __vimax3_s32(d1, v1, kTieNone);
                } else {
                    const int p0 = __viaddmax_s32(D0, s[q][0], kTieNone);
                    const int p1 = __viaddmax_s32(D1, s[q][1], kTieNone);
                    t0 = __viaddmax_s32(u0, gu, p0);
                    t1 = __viaddmax_s32(u1, gu, p1);
                }
                dg = hl[q];                              // the row above the next row at column c0 - 1
                const int k0 = __viaddmax_s32(L0, gl, t0);
                const int h0 = (k0 & ~15) | h7r;
                if (q == kR - 1) nA0 = __shfl_up_sync(0xffffffffu, h0, 1);
                const int k1 = __viaddmax_s32(h0, gl, t1);
                const int h1 = (k1 & ~15) | h7r;
                if (q == kR - 1) nA1 = __shfl_up_sync(0xffffffffu, h1, 1);
                hA0[q] = h0; hA1[q] = h1; kA0[q] = k0; kA1[q] = k1;
                u0 = h0; u1 = h1;
            }
        }
        // ---------------- second half: columns c0 + 2, c0 + 3
        const bool l0 = (lane == 0);
        const int b2 = l0 ? (PROF ? ((v.x & ~15) | h7r) : (v.x & ~15)) : nA0;
        const int b3 = l0 ? v.y : nA1;
        int o0 = 0, o1 = 0, o2 = 0, o3 = 0;
        {
            int u2 = b2, u3 = b3, dg = A1;               // the row above at column c0 + 1
#pragma unroll
            for (int q = 0; q < kR; ++q) {
                int D2 = dg, D3 = u2, L2 = hA1[q];
                if (MODE & 1) {
                    D2 = (c0 + 2 <= 0) ? kSNeg : D2; D3 = (c0 + 3 <= 0) ? kSNeg : D3;
                    L2 = (c0 + 2 == 0) ? lb[q] : L2;
                }
                int t2, t3;
                if constexpr (PROF) {
                    const int d2 = __dp4a(sb[q], 16 << 16, D2), d3 = __dp4a(sb[q], 16 << 24, D3);
                    const int v2 = addf(u2, gu), v3 = addf(u3, gu);
                    t2 = __vimax3_s32(d2, v2, kTieNone);
                    t3 = __vimax3_s32(d3, v3, kTieNone);
                } else {
                    const int p2 = __viaddmax_s32(D2, s[q][2], kTieNone);
                    const int p3 = __viaddmax_s32(D3, s[q][3], kTieNone);
                    t2 = __viaddmax_s32(u2, gu, p2);
                    t3 = __viaddmax_s32(u3, gu, p3);
                }
                dg = hA1[q];
                const int k2 = __viaddmax_s32(L2, gl, t2);
                const int h2 = (k2 & ~15) | h7r;
                if (q == kR - 1) nB2 = __shfl_up_sync(0xffffffffu, h2, 1);
                const int k3 = __viaddmax_s32(h2, gl, t3);
                const int h3 = (k3 & ~15) | h7r;
                if (q == kR - 1) nB3 = __shfl_up_sync(0xffffffffu, h3, 1);
                hl[q] = h3;
                const int k0 = kA0[q], k1 = kA1[q], h0 = hA0[q], h1 = hA1[q];
                if (STORE && SWB_KMAXC) {
                    int e0 = k0, e1 = k1, e2 = k2, e3 = k3;
                    if (MODE & 1) { e0 = (c0 == 0) ? 0 : e0; e2 = (c0 + 2 == 0) ? 0 : e2; }       // column 0 (columns < 0 hold NONE)
                    if (MODE & 2) {
                        e0 = (c0 <= mcols) ? e0 : 0;
                        e1 = (c0 + 1 <= mcols) ? e1 : 0;
                        e2 = (c0 + 2 <= mcols) ? e2 : 0;
                        e3 = (c0 + 3 <= mcols) ? e3 : 0;
                    }
                    kmax = __vimax3_s32(__vimax3_s32(kmax, e0, e1), e2, e3);
                }
                if (STORE) {
                    if (q == 0) sts_int4<0>(sa, k0, k1, k2, k3);
                    if (q == 1) sts_int4<kRowInts * 4>(sa, k0, k1, k2, k3);
                    if (q == 2) sts_int4<2 * kRowInts * 4>(sa, k0, k1, k2, k3);
                    if (q == 3) sts_int4<3 * kRowInts * 4>(sa, k0, k1, k2, k3);
                } else {
                    // score only: see step()
                    int v0 = addf(h0, 3 - kH7), v1 = addf(h1, 2 - kH7), v2 = addf(h2, 1 - kH7), v3 = addf(h3, 0 - kH7);
                    if (MODE != 0) {
                        v0 = (c0 >= 1 && c0 <= mcols) ? v0 : -1;
                        v1 = (c0 + 1 >= 1 && c0 + 1 <= mcols) ? v1 : -1;
                        v2 = (c0 + 2 >= 1 && c0 + 2 <= mcols) ? v2 : -1;
                        v3 = (c0 + 3 >= 1 && c0 + 3 <= mcols) ? v3 : -1;
                    }
                    const int bm = max(__vimax3_s32(v0, v1, v2), v3);
                    const bool upd = bm > rmax[q];
                    rbest[q] = upd ? bm : rbest[q];
                    rmax[q] = upd ? (bm | 15) : rmax[q];
                    rcol[q] = upd ? t : rcol[q];
                }
                u2 = h2; u3 = h3;
                if (q == kR - 1) { o0 = h0; o1 = h1; o2 = h2; o3 = h3; }
            }
        }
        dgp = b3;                                        // the row above at column c0 + 3 = before the next step's first column
        if (STORE) sa = ((sa + 16u) & (unsigned)(KT * 16 - 1)) | sa_base;
        // ---------------- hand my last row to the next strip (lane 31 only) ----------------
        {
            const int4 o = make_int4(PROF ? (o0 ^ (LAST ? otag_w : otag)) : (o0 | (LAST ? otag_w : otag)), o1, o2, o3);
            if (LAST) sts_volatile_int4_if<0>(out_w, o, out_ring);
            else      sts_volatile_int4_if<16 * I>(out_g, o, out_ring);
            st_cg_int4_if<16 * I>(gout, o, out_glob);
        }
        // ---------------- scores of the next step ----------------
        if constexpr (PROF) lookup(next_word);
        else                scores(next_word);
        // ---------------- the row above the first half of the next step ----------------
        A0 = l0 ? v.z : nB2;
        A1 = l0 ? v.w : nB3;
    }
};

// Flags live in shared memory and are passed as 32-bit shared addresses.  Ordering between
// the staged data and its flag relies on the in-order shared-memory pipeline of one warp
// (data STS, __syncwarp, flag STS; flag LDS, data LDS): a MEMBAR here waits for the writer's
// outstanding GLOBAL stores as well and cost ~2500 clk per round.
template <int KT, bool STORE, bool PROF>
__device__ __forceinline__ void compute_strip(const FillParams& p, Strip<KT, STORE, PROF>& S, const unsigned* aw,
                                              const unsigned staged, const unsigned drained,
                                              const unsigned consumed_in, const unsigned consumed_out,
                                              const bool ring_consumer, const long long strip)
{
    const int lane = S.lane;
    trace_stamp(p, strip, 0, lane);
    // packed characters / selector words: this group's, the next group's, and the one after (loaded two groups = ~4000
    // clk ahead: under the writers' store traffic a global load takes far longer than an idle L2 hit, and one group of
    // distance left the warp waiting 300-400 clk per group for these eight words)
    constexpr bool kPf2 = kHS;     // (the two-row geometry runs two CTAs per SM and has no registers for the third set)
    unsigned cur[kGroup + 1], nxt[kGroup], nx2[kGroup];
#pragma unroll
    for (int i = 0; i < kGroup; ++i) { cur[i] = __ldg(aw + i); nxt[i] = __ldg(aw + kGroup + i); }

    // first input entry (entry x -> ring index (x+32)&63, epoch 0 -> tag 1); all lanes poll.  Half skew: entry -16 holds
    // the row above columns -2 .. 1; its .zw is the row above the first half of step 0.
    if (S.has_in) {
        int4 v = lds_volatile_int4<16 * (32 + kFirstEntry)>(S.ring_in);
        while ((v.x & 3) != 1) { if (SWB_X_GATESLEEP > 0) __nanosleep(SWB_X_GATESLEEP); v = lds_volatile_int4<16 * (32 + kFirstEntry)>(S.ring_in); }
        if (lane == 0) {
            if (kHS) { S.A0 = v.z; S.A1 = v.w; S.dgp = v.y; }
            else     { S.A0 = (v.x & ~15) | S.kH7; S.A1 = v.y; S.A2 = v.z; S.A3 = v.w; }
        }
    }
    trace_stamp(p, strip, 1, lane);
#ifdef SWB_X_CLKTRACE
    const long long clk_gate = clock64();
#endif
    // (with the score look-up a NUL byte in b needs no care: positions outside a carry the pad code, which matches no row)
    const bool forced = (STORE && p.left_in != nullptr) || (!PROF && opaque(*p.nul_flag) != 0);
    if constexpr (PROF)    S.lookup(cur[0]);
    else if (forced && !kHS) S.head_fix(cur[0], -lane);
    else                   S.scores(cur[0]);

    const int gtail = (p.in_last - 7 - kInOfs) >> 3;      // groups g <= gtail: every entry their steps load exists (<= in_last)
    int known_drained = 0, known_consumed = 0;
#ifdef SWB_X_GROUPTRACE
    long long dbg_pre = 0, dbg_steps = 0, dbg_post = 0, dbg_n = 0, dbg_e_steps = 0, dbg_e_other = 0, dbg_drain = 0, dbg_cons = 0;
#endif
    for (int g = 0; g < p.ngroups; ++g) {
#ifndef SWB_X_GROUPTRACE
        if (SWB_TRACE && p.trace != nullptr && (g == 4 || g == 8)) trace_stamp(p, strip, g == 4 ? 2 : 3, lane);
#endif
        const int t0 = g * kGroup;
#ifdef SWB_X_GROUPTRACE
        const long long gc0 = clock64();
#endif
        // ---- staging ring space: the writers must have drained the slots this group overwrites
        //      (the flags only grow: the last values read are kept as credits, so a strip whose writers and
        //      consumer keep up reads them once every few groups instead of three times per group)
#ifdef SWB_X_GROUPTRACE
        const long long gd0 = clock64();
#endif
        if (STORE && g - stage_slack(KT) > known_drained) {
            int d;
            do {
                d = lds_volatile_int(drained);
#pragma unroll
                for (int k = 1; k < kWriters; ++k) d = min(d, lds_volatile_int(drained + 4u * k));
            } while (d < g - stage_slack(KT));
            known_drained = d;
        }
#ifdef SWB_X_GROUPTRACE
        const long long gd1 = clock64();
#endif
        // ---- hand-off ring space (blocks up to t0+7-31 are written in this group)
        // (this group writes entries up to t0 + 7 - 31 over entries kRing older; at its group start c the consumer has
        //  consumed the entries up to c + kInOfs - 1)
        if (ring_consumer && t0 - (kRing + 15 + kInOfs) > known_consumed) {
            int c;
            do { c = lds_volatile_int(consumed_out); } while (c < t0 - (kRing + 15 + kInOfs));
            known_consumed = c;
        }
#ifdef SWB_X_GROUPTRACE
        if (g >= SWB_X_GT_LO && g < SWB_X_GT_HI) { const long long gd2 = clock64(); dbg_drain += gd1 - gd0; dbg_cons += gd2 - gd1; }
#endif
        sts_volatile_int_if(consumed_in, t0, S.has_in & (lane == 0 ? 1 : 0));
        // ---- sequence words of the next group
#pragma unroll
        for (int i = 0; i < kGroup; ++i) nx2[i] = __ldg(aw + t0 + (kPf2 ? 2 : 1) * kGroup + i);
        cur[kGroup] = kPf2 ? nxt[0] : nx2[0];

        // consumer side: block t -> ring index (t+32)&63, epoch ((t+32)>>6)&1
        // (step t0 + I loads entry t0 + I + kInOfs = ring index in_g + I + 1)
        constexpr int kRS = 32 + kInOfs - 1;
        const unsigned in_g  = S.ring_in + 16u * (unsigned)((t0 + kRS) & (kRing - 1));
        const unsigned in_w  = S.ring_in + 16u * (unsigned)((t0 + kRS + 8) & (kRing - 1));
        const int   want     = 1 + (((t0 + kRS) >> kRingLog) & 1);
        const int   want_w   = 1 + (((t0 + kRS + 8) >> kRingLog) & 1);
        // producer side: block t-31 -> ring index (t+1)&63, epoch ((t+1)>>6)&1 (= its (j+32) form)
        const unsigned out_g = S.ring_out + 16u * (unsigned)((t0 & (kRing - 1)) + 1);
        const unsigned out_w = S.ring_out + 16u * (unsigned)((t0 + 8) & (kRing - 1));
        const int   otag     = (1 + ((t0 >> kRingLog) & 1)) ^ S.kH7;
        const int   otag_w   = (1 + (((t0 + 8) >> kRingLog) & 1)) ^ S.kH7;

#ifdef SWB_X_GROUPTRACE
        const long long gc1 = clock64();
#endif
        constexpr int kW = (KT == 64 && STORE) ? SWB_WAIT_STEPS_SINGLE : kWaitSteps;
        // The strip above is polled once per run of kW steps, for the LAST block the run needs (the
        // producer publishes its blocks in order); inside a run there is no branch.  This costs kW
        // steps of extra lag behind the producer and saves a third of every step.
#define SWB_WAIT(H)                                                                                       \
        if (S.has_in) {                                                                                   \
            const int xb = min(t0 + kW * ((H) + 1) + kInOfs - 1, p.in_last);                      \
            if (xb > t0 + kW * (H) + kInOfs - 1) wait_block(S.ring_in, xb);                       \
        }
        // the same for a group that ends before jmax: block t0+4 sits four entries after block t0 (no wrap: t0 is a
        // multiple of 8) in the same epoch, block t0+8 is the entry the last step polls anyway
#define SWB_WAITF(H)                                                                                      \
        if (S.has_in) {                                                                                   \
            if (kW * ((H) + 1) < 8) { while ((lds_volatile_int(in_g + 16u * (unsigned)(kW * ((H) + 1))) & 3) != want) { } }  \
            else                    { while ((lds_volatile_int(in_w) & 3) != want_w) { } }                \
        }
#define SWB_STEP(E, I) { if constexpr (kHS) S.template step_hs<E, I>(t0 + I, cur[I + 1], in_g, in_w, want, want_w, out_g, out_w, otag, otag_w); \
                         else               S.template step<E, I>(t0 + I, cur[I + 1], in_g, in_w, want, want_w, out_g, out_w, otag, otag_w); }
        // forced = b holds a NUL byte, or column-strip mode (boundary injection): head and tail steps differ.
        // Otherwise a full fill runs the interior step everywhere; score only still masks the columns past m.
#define SWB_GROUP(M, W) { W(0) SWB_STEP(M, 0); SWB_STEP(M, 1);                                        \
                          if (kW == 2) { W(1) }                                               \
                          SWB_STEP(M, 2); SWB_STEP(M, 3);                                     \
                          if (kW == 4) { W(1) } else if (kW == 2) { W(2) }                    \
                          SWB_STEP(M, 4); SWB_STEP(M, 5);                                     \
                          if (kW == 2) { W(3) }                                               \
                          SWB_STEP(M, 6); SWB_STEP(M, 7); }
        // (only in the single-pair full-fill instantiation: the others are compiled for two CTAs per SM and have no
        // registers to spare for a fifth copy of the group)
        if ((KT == 64 && STORE) && g <= gtail && !(forced && g < 4)) SWB_GROUP(0, SWB_WAITF)
        else {
            const bool head = forced && g < 4, tail = (forced || !STORE) && g > gtail;
            if (!head && !tail) SWB_GROUP(0, SWB_WAIT)
            else if (head && !tail) SWB_GROUP(1, SWB_WAIT)
            else if (tail && !head) SWB_GROUP(2, SWB_WAIT)
            else SWB_GROUP(3, SWB_WAIT)
        }
#undef SWB_GROUP
#undef SWB_STEP
#undef SWB_WAIT
#undef SWB_WAITF
#ifdef SWB_X_GROUPTRACE
        const long long gc2 = clock64();
#endif
        S.gout += kGroup;
#pragma unroll
        for (int i = 0; i < kGroup; ++i) { cur[i] = kPf2 ? nxt[i] : nx2[i]; nxt[i] = nx2[i]; }
        // ---- publish the staged group to the writers
        __syncwarp();
        if (STORE) sts_volatile_int_if(staged, g + 1, lane == 0 ? 1 : 0);
#ifdef SWB_X_GROUPTRACE
        { const long long gc3 = clock64();
          if (g >= SWB_X_GT_LO && g < SWB_X_GT_HI) { dbg_pre += gc1 - gc0; dbg_steps += gc2 - gc1; dbg_post += gc3 - gc2; ++dbg_n; }
          else if (g < 4) { dbg_e_steps += gc2 - gc1; dbg_e_other += (gc1 - gc0) + (gc3 - gc2); } }
#endif
    }
#ifdef SWB_X_GROUPTRACE
    if (p.trace && lane == 0) {
        p.trace[strip * 8 + 2] = dbg_pre; p.trace[strip * 8 + 3] = dbg_steps; p.trace[strip * 8 + 4] = dbg_post;
        p.trace[strip * 8 + 5] = dbg_n; p.trace[strip * 8 + 6] = dbg_e_steps; p.trace[strip * 8 + 7] = ((dbg_drain >> 4) << 32) | ((dbg_cons >> 4) & 0xffffffffll); (void)dbg_e_other;
    }
    return;
#endif
    trace_stamp(p, strip, 4, lane);
#ifdef SWB_X_CLKTRACE
    if (p.trace && lane == 0) { p.trace[strip * 8 + 6] = clock64() - clk_gate; }
#endif
}

// ---------------------------------------------------------------------------------
// writer warp: drains 32 rows of the staging ring of one strip into H and P
//   sub = which kWRows rows of the strip (0 .. kWriters-1); table entries / row loops use the first kWRows lanes' rows
// ---------------------------------------------------------------------------------
template <int KT>
__device__ __forceinline__ void writer_strip(const FillParams& p, const long long r0, const int sub, const int lane,
                                             const int* stage /* this strip: [32*kR][4*KT] */,
                                             int4* rowtab,
                                             volatile int* staged, volatile int* drained, const long long strip)
{
    // strip row rho = 32*sub + lane is computed by compute lane cl = rho / kR.
    // per-row constants: round r flushes, for this row, the 32 columns 32r-E .. 32r-E+31 whose
    // first element sits on a 128-byte line of H (and P); E is the smallest such offset for
    // which lane cl has finished those columns by the end of compute group r
    constexpr int kRowInts = 4 * KT;
    const int rho = kWRows * sub + (lane & (kWRows - 1));
    const int cl  = rho / kR;
    const long long row = r0 + rho;
    const bool myrow = lane < kWRows && row <= p.n;
    // (the phase of the row's first element within its 128-byte line counts the base address too: pair k of a batch
    //  starts pair_stride ints after pair 0, a 16-byte aligned but in general not a 128-byte aligned address -- with the
    //  phase taken from row * pitch alone every "line" store of such a pair straddled two lines)
    const int ph = (int)((((unsigned long long)p.H >> 2) + (unsigned long long)(row * p.pitch)) & 31);
    const int d  = (kSkewCols * cl + 31 - ph) >> 5;               // (lane cl has finished column 32r + 31 - kSkewCols*cl by the end of group r)
    const int E  = 32 * d + ph;
    const long long G0 = row * p.pitch - E;                      // multiple of 32
    const int F  = ((4 + kSkewCols) * cl - E) & (kRowInts - 1);   // ring index of column c is (c + (4 + kSkewCols)*cl) mod kRowInts
    const unsigned long long hb = (unsigned long long)(p.H + G0);
    if (lane < kWRows) rowtab[lane] = make_int4((int)(unsigned)hb, (int)(unsigned)(hb >> 32), F, E);
    // the same for the fast path: byte offset of the row's segment base from the strip's first row (fits 32 bits:
    // kStripRows * pitch * 4 < 2^32 is checked by the host) and F
    int2* rowoff = reinterpret_cast<int2*>(rowtab + 32);
    const unsigned long long hbase = (unsigned long long)p.H + 4ull * (unsigned long long)(r0 * p.pitch - 1024);     // (E < 1024)
    if (lane < kWRows) rowoff[lane] = make_int2((int)(unsigned)(4 * (G0 - r0 * p.pitch + 1024)), F);
    const unsigned one = (unsigned)opaque(1);
    const unsigned rowmask = __ballot_sync(0xffffffffu, myrow);
    const int nvalid = __popc(rowmask);
    const int Emax = __reduce_max_sync(0xffffffffu, myrow ? E : 0);
    const int Emin = __reduce_min_sync(0xffffffffu, myrow ? E : 0x7fffffff);
    __syncwarp();
    const int* mystage = stage + (size_t)kWRows * sub * kRowInts;

    int mx = 0;                                                   // largest key of my rows in columns 1..m
    const int rounds = p.ngroups + kDrainRounds;
    const int m = (int)p.m;
    const long long pdelta = (long long)(p.P - p.H);              // P[i] sits pdelta ints after H[i]
#ifdef SWB_X_WRITERTRACE
    long long dw_wait = 0, dw_work = 0, dw_n = 0;
#endif
    for (int r = 0; r < rounds; ++r) {
#ifdef SWB_X_WRITERTRACE
        const long long wc0 = clock64();
#endif
        const int need = min(r + 1, p.ngroups);
        if (*staged < need) {
            int spins = 0;
            while (*staged < need) { if (++spins > 8) __nanosleep(SWB_X_WRITERSLEEP); }
        }
        asm volatile("" ::: "memory");
#ifdef SWB_X_WRITERTRACE
        const long long wc1 = clock64();
#endif
        const int v = 32 * r + lane;
        // (the valid rows of a strip are its first nvalid ones: the last strip of a pair may be partial)
        const bool interior = (32 * r - Emax >= (p.left_in != nullptr ? 1 : 0)) && (32 * r + 31 - Emin <= m - (p.right_out != nullptr ? 1 : 0));
#ifdef SWB_X_NOWRITER
        if (interior && nvalid == kWRows) {
        } else
#endif
        if (interior && nvalid == kWRows) {
            // batches of 8 rows: all table and data loads first, then the 16 stores.  Per row the ALU pipe
            // sees the ring index (2 ops) and the unpacking (2 ops); the two addresses are one IMAD.WIDE each
            // (row offset in bytes * 1 + the 64-bit address of this lane's column in the strip's first row)
#ifndef SWB_WRITER_DEPTH
#define SWB_WRITER_DEPTH 4
#endif
#ifndef SWB_WRITER_PAIRS
#define SWB_WRITER_PAIRS 1
#endif
#if SWB_WRITER_PAIRS
            // Two rows per store instruction: a half-warp covers the 32 columns of one row with 8-byte stores (every
            // row's window starts on a 128-byte line, so two adjacent columns are 8-byte aligned), lanes 0-15 row 2i,
            // lanes 16-31 row 2i+1.  Same bytes, same shared-memory loads, half the STG instructions in the SM's
            // load/store pipeline (which the compute warps' shuffles share).
            const int half = lane >> 4;
            const int vv = 32 * r + 2 * (lane & 15);                  // my two columns of the window: vv, vv + 1
            const unsigned long long hcol = hbase + 4ull * (unsigned)vv, pcol = hcol + 4ull * (unsigned long long)pdelta;
            constexpr int D = SWB_WRITER_DEPTH;
            int ka[D], kb[D]; unsigned off[D];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const int row = 2 * i + half;
                const int2 tb = rowoff[row];
                off[i] = (unsigned)tb.x;
                ka[i] = mystage[row * kRowInts + ((vv + tb.y) & (kRowInts - 1))];
                kb[i] = mystage[row * kRowInts + ((vv + 1 + tb.y) & (kRowInts - 1))];
            }
#pragma unroll
            for (int i = 0; i < kWRows / 2; ++i) {
                const int xa = ka[i % D], xb = kb[i % D]; const unsigned oo = off[i % D];
                if (i + D < kWRows / 2) {
                    const int row = 2 * (i + D) + half;
                    const int2 tb = rowoff[row];
                    off[i % D] = (unsigned)tb.x;
                    ka[i % D] = mystage[row * kRowInts + ((vv + tb.y) & (kRowInts - 1))];
                    kb[i % D] = mystage[row * kRowInts + ((vv + 1 + tb.y) & (kRowInts - 1))];
                }
#ifndef SWB_X_NOSTG
                __stcs(reinterpret_cast<int2*>(mad_wide(oo, one, hcol)), make_int2(xa >> 4, xb >> 4));
                __stcs(reinterpret_cast<int2*>(mad_wide(oo, one, pcol)), make_int2(xa & 3, xb & 3));
                if (!SWB_KMAXC) mx = max(mx, max(xa, xb));
#else
                if (xa == 0x7ffffff1) __stcs(reinterpret_cast<int2*>(mad_wide(oo, one, hcol)), make_int2(xa >> 4, xb >> 4));
#endif
            }
#else
            const unsigned long long hcol = hbase + 4ull * (unsigned)v, pcol = hcol + 4ull * (unsigned long long)pdelta;
            // rolling pipeline over the rows: the loads of row i+D are issued before the two stores of row i, so the
            // stores leave the SM as a steady trickle.  (Batches of 8 rows = bursts of 16 STGs: the shuffles and
            // shared-memory accesses on the compute warps' chain queue behind them in the SM's load/store pipeline.)
            constexpr int D = SWB_WRITER_DEPTH;
            int k[D]; unsigned off[D];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const int2 tb = rowoff[i];
                off[i] = (unsigned)tb.x;
                k[i] = mystage[i * kRowInts + ((v + tb.y) & (kRowInts - 1))];
            }
#pragma unroll
            for (int i = 0; i < kWRows; ++i) {
                const int kk = k[i % D]; const unsigned oo = off[i % D];
                if (i + D < kWRows) {
                    const int2 tb = rowoff[i + D];
                    off[i % D] = (unsigned)tb.x;
                    k[i % D] = mystage[(i + D) * kRowInts + ((v + tb.y) & (kRowInts - 1))];
                }
#ifndef SWB_X_NOSTG
                SWB_ST(reinterpret_cast<int32_t*>(mad_wide(oo, one, hcol)), kk >> 4);
                SWB_ST(reinterpret_cast<int32_t*>(mad_wide(oo, one, pcol)), kk & 3);
                if (!SWB_KMAXC) mx = max(mx, kk);
#else
                if (kk == 0x7ffffff1) __stcs(reinterpret_cast<int32_t*>(mad_wide(oo, one, hcol)), kk >> 4);
#endif
            }
#endif
        } else if (KT == 32 && p.left_in == nullptr && p.right_out == nullptr) {
            // Edge rounds (for some rows part of the round lies left of column 0 or right of column m) and the partial
            // last strip of a pair: batches of 8 rows with predicated stores.  For small matrices most rounds are edge
            // rounds (9 of 12 at 256 columns: the 65536 x 256x256 batch went from 302 to 355 GCUPS with this path).
            // Batch instantiation only: in the single-pair kernel the same path made the 45000x45000 fill 8 % SLOWER
            // (5.0 -> 5.4 ms; the row-by-row loop below spreads the first rounds' stores of a strip thinner).
#pragma unroll 1
            for (int l0 = 0; l0 < kWRows; l0 += 8) {
                int k[8]; int32_t* hp[8]; bool ok[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int4 tb = rowtab[l0 + i];                                    // (pointer, F, E)
                    hp[i] = reinterpret_cast<int32_t*>(((unsigned long long)(unsigned)tb.y << 32) | (unsigned)tb.x) + v;
                    ok[i] = ((rowmask >> (l0 + i)) & 1u) && (unsigned)(v - tb.w) <= (unsigned)m;
                    k[i] = mystage[(l0 + i) * kRowInts + ((v + tb.z) & (kRowInts - 1))];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (ok[i]) {
                        SWB_ST(hp[i], k[i] >> 4);
                        SWB_ST(hp[i] + pdelta, k[i] & 3);
                        if (!SWB_KMAXC) mx = max(mx, k[i]);          // (column 0 holds NONE = 8 >> 4 = 0)
                    }
                }
            }
        } else if (interior) {
            // column-strip mode, partial strip (the last one of a pair): row by row
#pragma unroll 2
            for (int l = 0; l < nvalid; ++l) {
                const int4 tb = rowtab[l];
                int32_t* hp = reinterpret_cast<int32_t*>(((unsigned long long)(unsigned)tb.y << 32) | (unsigned)tb.x) + v;
                const int k = mystage[l * kRowInts + ((v + tb.z) & (kRowInts - 1))];
                __stcs(hp, k >> 4);
                __stcs(hp + pdelta, k & 3);
                if (!SWB_KMAXC) mx = max(mx, k);
            }
        } else {
#pragma unroll 2
            for (int l = 0; l < kWRows; ++l) {
                const int4 tb = rowtab[l];
                const int idx = (v + tb.z) & (kRowInts - 1);
                const int c = v - tb.w;
                if (((rowmask >> l) & 1u) && c >= 0 && c <= m) {
                    const int k = mystage[l * kRowInts + idx];
                    int32_t* hp = reinterpret_cast<int32_t*>(((unsigned long long)(unsigned)tb.y << 32) | (unsigned)tb.x) + v;
                    __stcs(hp, k >> 4);
                    // column-strip mode: local column 0 belongs to the GPU on the left; its P holds the
                    // hand-off marker that ends this GPU's part of the backtrack
                    __stcs(hp + pdelta, (c == 0 && p.left_in != nullptr) ? kHandOff : (k & 3));
                    if (!SWB_KMAXC && c > 0) mx = max(mx, k);
                    // ... and my last column is the boundary column of the GPU on the right (P2P store over NVLink)
                    if (c == m && p.right_out != nullptr)
                        asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p.right_out + r0 + kWRows * sub + l), "r"(k >> 4) : "memory");
                }
            }
        }
        __syncwarp();
        asm volatile("" ::: "memory");
        if (lane == 0) *drained = r + 1;
#ifdef SWB_X_WRITERTRACE
        if (interior) { const long long wc2 = clock64(); dw_wait += wc1 - wc0; dw_work += wc2 - wc1; ++dw_n; }
#endif
    }
#ifdef SWB_X_WRITERTRACE
    if (p.trace && lane == 0 && sub == 0) { p.trace[strip * 8 + 5] = dw_wait; p.trace[strip * 8 + 6] = dw_work; p.trace[strip * 8 + 7] = dw_n; }
#endif
    if (p.right_flags != nullptr && nvalid > 0) {
        // column-strip mode: every lane's boundary stores are ordered before the flag the right GPU polls
        __threadfence_system();
        __syncwarp();
        if (lane == 0)
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p.right_flags + (r0 - 1) / kWRows + sub), "r"(p.epoch) : "memory");
    }
    if (!SWB_KMAXC) {
        // strip maximum (omp_smithW.c:384-387 needs only the arg-max; see argmax_kernel)
        const int hm = __reduce_max_sync(0xffffffffu, mx) >> 4;
        if (lane == 0 && hm > 0) { atomicMax(p.strip_max + strip, hm); atomicMax(p.gmax, hm); }
    }
#if !defined(SWB_X_GROUPTRACE) && !defined(SWB_X_WRITERTRACE) && !defined(SWB_X_CLKTRACE)
    trace_stamp(p, strip, 5 + (sub & 1), lane);
#endif
}

// ---------------------------------------------------------------------------------
// loader warp: copies the tagged last row of the band above from global memory (L2)
// into the first hand-off ring of this band (same slot / epoch tag arithmetic)
// ---------------------------------------------------------------------------------
// src = entry 0 of the boundary row; entries kFirstEntry .. last are copied.  An entry x may overwrite the ring slot of
// entry x - kRing once the consumer has used it: at its group start c it has consumed the entries up to c + kInOfs - 1.
__device__ __forceinline__ void loader_band(const int4* src, const int last, int4* ring, const int lane,
                                            volatile int* consumed)
{
    const int nblocks = last + 1;
    int base = kFirstEntry;
    int idle = 0;
    while (base < nblocks) {
        const int limit = min(nblocks, *consumed + kRing + kInOfs - 1);
        const int j = base + lane;
        bool ok = false;
        int4 v = make_int4(0, 0, 0, 0);
        if (j < limit) {
            v = ld_cg_int4(src + j);
            ok = (v.x & 3) == 1 + (((j + 32) >> kRingLog) & 1);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        const int lead = (mask == 0xffffffffu) ? 32 : (__ffs(~mask) - 1);
        // blocks enter the ring in order (the consumer polls only the last block of a run of steps)
        if (lane < lead) sts_volatile_int4(ring + ((j + 32) & (kRing - 1)), v);
        base += lead;
        // (polls without sleeping while the band above is producing: a sleep here is microseconds of lag on the fill's
        // critical chain; after a long silence the band above has not started yet and the loader backs off)
        if (lead == 0) { if (++idle > SWB_X_LOADERSPINS) __nanosleep(100); } else idle = 0;
    }
}

// threads per block for wpc strips per band: rows of 4 warps; the 4-wpc serving schedulers
// hold wpc*kWriters writers (none in score-only mode) + 1 loader
__host__ __device__ constexpr int fill_block_threads(int wpc, bool store)
{
    return 32 * 4 * (((store ? wpc * kWriters : 0) + 1 + (4 - wpc) - 1) / (4 - wpc));
}
__host__ __device__ constexpr size_t fill_smem_bytes(int wpc, int KT, bool store)
{
    return (size_t)wpc * ((store ? (size_t)kStripRows * 4 * KT * sizeof(int) + kWriters * 48 * sizeof(int4) : 0) +
                          kRing * sizeof(int4));
}

// ---------------------------------------------------------------------------------
// The fill kernel.  grid = number of bands (x pairs), block = fill_block_threads(wpc) threads
// (roles by warp id, see below): wpc compute warps, wpc*kWriters writers, one loader of the
// band boundary.  dynamic smem = fill_smem_bytes(wpc, KT, STORE).
// STORE = false is the score-only variant: no staging, no writers, per-row best cells instead.
// ---------------------------------------------------------------------------------
template <int KT, bool STORE, bool PROF>
__device__ __forceinline__ void fill_body(const FillParams& p_in)
{
    extern __shared__ __align__(1024) int4 smem4[];
    __shared__ int s_band;
    __shared__ int s_staged[kMaxWpc], s_drained[kMaxWpc * kWriters], s_consumed[kMaxWpc + 1];

    const int lane = threadIdx.x & 31;
    const int wid  = threadIdx.x >> 5;
    const int wpc  = p_in.wpc;
    // Warp roles by scheduler (a warp runs on SM sub-partition wid % 4).  A compute warp must
    // not share its scheduler with another busy warp (measured: 190 -> 335 clk per step), so
    // the compute warps take schedulers 0..wpc-1 (first row of warps) and the writers and the
    // loader are spread over the other schedulers; the remaining warp slots exit at once.
    // Several CTAs can share an SM (batches, score-only): consecutive bands rotate the scheduler
    // assignment so that their compute warps land on different schedulers.
    if (threadIdx.x == 0) s_band = atomicAdd(p_in.ticket, 1);
    if (threadIdx.x < kMaxWpc) s_staged[threadIdx.x] = 0;
    if (threadIdx.x < kMaxWpc * kWriters) s_drained[threadIdx.x] = 0;
    if (threadIdx.x <= kMaxWpc) s_consumed[threadIdx.x] = 0;
    __syncthreads();
    const int sched = ((wid & 3) + 4 - ((s_band * wpc) & 3)) & 3, wrow = wid >> 2;
    const int nserv = 4 - wpc;                                   // schedulers that serve writers / loader
    const int nwriters = STORE ? wpc * kWriters : 0;
    int role = -1;                                               // -1 idle, 0 compute, 1 writer, 2 loader
    int w = 0;                                                   // compute: strip in the band; writer: writer index
    if (sched < wpc) { if (wrow == 0) { role = 0; w = sched; } }
    else {
        const int slot = wrow * nserv + (sched - wpc);
        if (slot < nwriters) { role = 1; w = slot; }
        else if (slot == nwriters) role = 2;
    }

    int4* stage4  = smem4;                                       // [wpc][kStripRows][KT]   (STORE only)
    int4* rings   = stage4 + (STORE ? (size_t)wpc * kStripRows * KT : 0);   // [wpc][kRing]
    int4* rowtabs = rings + (size_t)wpc * kRing;                 // [wpc*kWriters][48]      (STORE only)

    // (tag 0 = not valid yet; the first strip of a pair reads its never-written ring as the zero row above row 1)
    for (int i = threadIdx.x; i < wpc * kRing; i += blockDim.x) rings[i] = make_int4(0, PROF ? kTieDiag : 0, PROF ? kTieDiag : 0, PROF ? kTieDiag : 0);
    __syncthreads();
    // Tickets are handed out in start order, so a waiting band's predecessor is resident or done.  Single pair: band by
    // band.  Batches: BAND-MAJOR over the pairs (band 0 of every pair, then band 1 of every pair ...): by the time band
    // b of a pair is scheduled, its band b-1 finished a whole wave of CTAs earlier and its boundary row is complete, so
    // no CTA of a batch ever occupies an SM while it waits for the strip above (pair-major order: 4 chained CTAs per
    // 256-row pair, each idling ~40 steps + an L2 round trip behind the previous one).
    // (65536 x 256x256: score only 7.40 -> 5.21 ms, full fill 11.36 -> 9.07 ms.  An earlier measurement had the full fill
    //  slower band-major (12.4 against 11.8 ms); it was taken while the writers' line phase ignored the base address of a
    //  pair, see writer_strip.)
    const bool band_major = p_in.npairs > 1;
    const long long pair = band_major ? (long long)(s_band % p_in.npairs) : (long long)(s_band / p_in.nbands);
    const int band = band_major ? (int)(s_band / p_in.npairs) : (s_band % p_in.nbands);
    FillParams p = p_in;
    p.a4 += pair * p.a4_stride; p.a4s += pair * p.a4_stride;
    p.b += pair * p.n;
    p.H += pair * p.pair_stride; p.P += pair * p.pair_stride;
    p.boundary += pair * (long long)(p.nbands - 1) * p.bstride;
    p.strip_max += pair * p.nstrips;
    p.gmax += pair;
    if (!STORE) p.row_best += pair * (p.n + 1);
    const long long band_r0 = 1 + (long long)band * wpc * kStripRows;

    if (role == 0) {
        // ------------------------------------------------ compute
        const long long r0 = band_r0 + (long long)kStripRows * w;
        if (r0 > p.n) return;
        Strip<KT, STORE, PROF> S;
        constexpr int kH7 = Strip<KT, STORE, PROF>::kH7;
        S.lane = lane;
#pragma unroll
        for (int q = 0; q < kR; ++q) {
            const long long row = r0 + kR * lane + q;
            S.b4[q] = (row <= p.n) ? (unsigned)p.b[row - 1] * 0x01010101u : 0u;
            S.inv[q] = (row <= p.n) ? 0u : 0x01010101u;
            S.hl[q] = kH7; S.rmax[q] = 15; S.rcol[q] = 0; S.rbest[q] = 0; S.kmax = 0;
            if constexpr (PROF) {
                // score bytes of this row's character against the codes 0..7: match for its own code, mismatch for the
                // others (the pad code and rows past n never match)
                const unsigned mm = (unsigned)p.mismatch8 & 0xffu, mt = (unsigned)p.match8 & 0xffu;
                const unsigned code = (row <= p.n) ? (unsigned)p.lmap[p.b[row - 1]] : (unsigned)kPadCode;
                unsigned lo = mm * 0x01010101u, hi = mm * 0x01010101u;
                if (code < 4) lo = (lo & ~(0xffu << (8 * code))) | (mt << (8 * code));
                else if (code < kPadCode) hi = (hi & ~(0xffu << (8 * (code - 4)))) | (mt << (8 * (code - 4)));
                S.tlo[q] = lo; S.thi[q] = hi;
            } else {
                S.tlo[q] = S.thi[q] = 0u;
            }
            S.sb[q] = 0;
        }
        S.one = opaque(1);
        S.h7r = opaque(kH7);
        // keep the scoring constants in registers: a shuffle result is opaque to ptxas, which
        // otherwise re-reads them from the constant bank at the head of every step, on the
        // dependency chain
        // (PROF: the H operands carry kH7 in their low bits, the gap constants take it off again)
        S.sm = opaque(p.s_match); S.sx = opaque(p.s_mismatch); S.gu = opaque(p.g_up - kH7); S.gl = opaque(p.g_left - kH7); S.g16 = opaque(p.g_left - kTieLeft);
        S.A0 = S.A1 = S.A2 = S.A3 = kH7; S.dgp = kH7;
        S.sa_base = (unsigned)__cvta_generic_to_shared(stage4 + ((size_t)w * kStripRows + (size_t)kR * lane) * KT);
        S.sa = S.sa_base + 16u * (unsigned)lane;                 // slot (t + lane) & (KT-1) at t = 0
        S.ring_in  = (unsigned)__cvta_generic_to_shared(rings + (size_t)w * kRing);
        S.ring_out = (unsigned)__cvta_generic_to_shared(rings + (size_t)(w + 1 < wpc ? w + 1 : w) * kRing);
        S.jmax = p.in_last;
        S.mcols = (int)p.m;
        {
            const int g16 = p.g_left - kTieLeft;                 // 16 * gap
            if (STORE && p.left_in != nullptr) {
                // column-strip mode: the boundary values of this strip come from the GPU on the left, whose
                // writer warps publish them in pieces of 32 rows
#pragma unroll
                for (int k = 0; k < kWriters; ++k) {
                    if (r0 + kWRows * k > p.n) break;
                    const int* f = p.left_flags + (r0 - 1) / kWRows + k;
                    int v;
                    unsigned spins = 0;
                    do {
                        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                        if (v != p.epoch) {
                            __nanosleep(200);
                            // bounded: if the left neighbour never runs (it failed, or the caller enqueued the strips in
                            // the wrong order on one device) the launch ends with an error instead of hanging the GPU
                            if (++spins > kGateSpinLimit) __trap();
                        }
                    } while (v != p.epoch);
                }
#pragma unroll
                for (int q = 0; q < kR; ++q) {
                    const long long row = r0 + kR * lane + q;
                    int hv = 0;
                    if (row <= p.n) asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(hv) : "l"(p.left_in + row) : "memory");
                    S.lb[q] = 16 * hv - g16 + kH7;
                }
            } else {
#pragma unroll
                for (int q = 0; q < kR; ++q) S.lb[q] = -g16 + kH7;
            }
        }
        S.has_in = opaque(r0 > 1 ? 1 : 0);
        const bool next_row = (r0 + kStripRows <= p.n);          // a strip below exists
        const bool ring_consumer = next_row && (w + 1 < wpc);
        S.out_ring = (ring_consumer && lane == 31) ? 1 : 0;
        S.out_glob = (next_row && (w + 1 == wpc) && lane == 31) ? 1 : 0;
        // block j = t - lane of step t goes to gout[j]: base of the group's first step, the
        // step index is an immediate (only lane 31's copy is ever dereferenced, from t = 31 on)
        S.gout = p.boundary + (size_t)(next_row && (w + 1 == wpc) ? band : 0) * p.bstride + kBoundaryPad - lane;
        // aw[t] = characters / selectors of my four columns of step t (half skew: odd lanes read the copy that is
        // shifted by two columns)
        const unsigned* aw = kHS ? ((lane & 1) ? p.a4s : p.a4) + kAPad - ((lane + 1) >> 1) : p.a4 + kAPad - lane;
        const long long strip = (r0 - 1) / kStripRows;
        compute_strip<KT, STORE, PROF>(p, S, aw, (unsigned)__cvta_generic_to_shared(s_staged + w),
                      (unsigned)__cvta_generic_to_shared(s_drained + w * kWriters),
                      (unsigned)__cvta_generic_to_shared(s_consumed + w),
                      (unsigned)__cvta_generic_to_shared(s_consumed + w + 1), ring_consumer, strip + pair * p.nstrips);
        if (STORE && SWB_KMAXC) {
            // strip maximum (omp_smithW.c:384-387 needs only the arg-max; see argmax_kernel).  The rows past n of a
            // partial strip never match (inv), so their cells stay below the cells above them.
            const int mx = __reduce_max_sync(0xffffffffu, S.kmax) >> 4;
            if (lane == 0 && mx > 0) { atomicMax(p.strip_max + strip, mx); atomicMax(p.gmax, mx); }
        }
        if (!STORE) {
            // per-row best cells and the strip / global maxima (maxPos is reduced from them)
            int mx = 0;
#pragma unroll
            for (int q = 0; q < kR; ++q) {
                const long long row = r0 + kR * lane + q;
                const int col = 4 * S.rcol[q] - kSkewCols * lane + 3 - (S.rbest[q] & 15);      // step -> my columns of it, low bits -> which
                if (row <= p.n)
                    p.row_best[row] = ((unsigned long long)(unsigned)(S.rbest[q] >> 4) << 32) | (0xffffffffu - (unsigned)col);
                if (row <= p.n) mx = max(mx, S.rbest[q] >> 4);
            }
            mx = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0 && mx > 0) { atomicMax(p.strip_max + strip, mx); atomicMax(p.gmax, mx); }
        }
    } else if (role == 1) {
        // ------------------------------------------------ writer
        const int wi = w;                                        // writer index in the CTA
        const int cw = wi / kWriters, sub = wi % kWriters;
        const long long r0 = band_r0 + (long long)kStripRows * cw;
        if (r0 > p.n) return;                                    // (a writer without valid rows still runs: it owns a drained flag)
        writer_strip<KT>(p, r0, sub, lane, reinterpret_cast<const int*>(stage4 + (size_t)cw * kStripRows * KT),
                     rowtabs + (size_t)wi * 48, s_staged + cw, s_drained + wi,
                     (r0 - 1) / kStripRows);
    } else if (role == 2) {
        // ------------------------------------------------ loader
        if (band == 0 || band_r0 > p.n) return;
        loader_band(p.boundary + (size_t)(band - 1) * p.bstride + kBoundaryPad, p.in_last, rings, lane, s_consumed);
    }
}

// The kernels.  The alphabet of b (counted on the device by selector_kernel, so that the call stays asynchronous for
// device-resident sequences) decides which form of the cell arithmetic runs, the same in every thread of the grid.
//  * fill_kernel<KT, STORE>: ONE launch holds both forms and branches (SWB_MERGED_FORMS, the batch geometry: two
//    launches, one of which returns at once, cost the 65536-pair batch 0.14 ms of 9.07 -- 262 144 empty CTAs -- and the
//    score-only batch 0.16 of 4.92 ms).
//  * fill_kernel_form<KT, STORE, PROF>: one kernel per form, both launched, the one that does not apply returns at
//    once (the single-pair geometries: there the merged kernel measured 1-4 % SLOWER on the chain-bound configurations
//    -- score-only 45000 x 45000 3.64 -> 3.72 ms, 2 000 000 x 1000 51.8 -> 53.8 ms -- and an empty launch of 235 CTAs
//    costs 3 us).
#ifndef SWB_MERGED_FORMS
#define SWB_MERGED_FORMS 0
#endif
constexpr bool kMergedForms = SWB_MERGED_FORMS != 0;

template <int KT, bool STORE>
__global__ void __launch_bounds__(fill_block_threads(kMaxWpc, true), ((KT == 64 && STORE) || kR > 2) ? 1 : 2)
fill_kernel(const FillParams p_in)
{
#ifdef SWB_X_FORCE_COMPARE                                     // developer build: always the character-compare form
    fill_body<KT, STORE, false>(p_in);
#else
    if (p_in.prof_ok != 0 && *p_in.nletters <= kMaxLetters) fill_body<KT, STORE, true>(p_in);
    else                                                    fill_body<KT, STORE, false>(p_in);
#endif
}

template <int KT, bool STORE, bool PROF>
__global__ void __launch_bounds__(fill_block_threads(kMaxWpc, true), ((KT == 64 && STORE) || kR > 2) ? 1 : 2)
fill_kernel_form(const FillParams p_in)
{
#ifdef SWB_X_FORCE_COMPARE
    if (PROF) return;
#else
    if ((p_in.prof_ok != 0 && *p_in.nletters <= kMaxLetters) != PROF) return;
#endif
    fill_body<KT, STORE, PROF>(p_in);
}

// score-only: maxPos from the per-row best cells -- among the rows whose best score equals the
// global maximum, the cell with the smallest i+j, then the largest i (the row's first column
// with that score is its earliest anti-diagonal).  One block per pair.
__global__ void rowbest_argmax_kernel(const unsigned long long* __restrict__ row_best, long long n, long long pitch,
                                      const int* __restrict__ gmax, long long* maxPos, int32_t* maxScore)
{
    const long long pair = blockIdx.x;
    row_best += pair * (n + 1);
    const int g = gmax[pair];
    __shared__ unsigned long long s_key;
    if (threadIdx.x == 0) s_key = ~0ull;
    __syncthreads();
    if (g > 0) {
        unsigned long long best = ~0ull;
        for (long long r = 1 + threadIdx.x; r <= n; r += blockDim.x) {
            const unsigned long long v = row_best[r];
            if ((int)(v >> 32) == g) {
                const long long col = (long long)(0xffffffffu - (unsigned)(v & 0xffffffffu));
                const unsigned long long k = ((unsigned long long)(r + col) << 32) | (unsigned long long)(0xffffffffu - (unsigned)r);
                best = k < best ? k : best;
            }
        }
        atomicMin(&s_key, best);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long pos = 0;
        if (g > 0) {
            const long long r = (long long)(0xffffffffu - (unsigned)(s_key & 0xffffffffu));
            pos = r * pitch + ((long long)(s_key >> 32) - r);
        }
        if (maxPos) maxPos[pair] = pos;
        if (maxScore) maxScore[pair] = g;
    }
}

// ---------------------------------------------------------------------------------
// maxPos with the reference's tie-break: among the cells with H == global max, the one
// with the smallest i+j, then the largest i (omp_smithW.c:203-215,282-291,384-387).
// Only rows of strips whose maximum equals the global maximum are scanned.
// ---------------------------------------------------------------------------------
__global__ void argmax_kernel(const int32_t* __restrict__ H, long long pitch, long long pair_stride, long long m, long long n,
                              const int* __restrict__ strip_max, const int* __restrict__ gmax,
                              unsigned long long* key)
{
    const long long pair = blockIdx.x;                             // grid = (pairs, blocks per pair)
    const long long nstrips = (n + kStripRows - 1) / kStripRows;
    H += pair * pair_stride; strip_max += pair * nstrips; key += pair;
    const int g = gmax[pair];
    if (g <= 0) return;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.y * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.y * blockDim.x) >> 5;
    constexpr int kChunk = 1024;                                   // columns per work item
    const long long nchunks = (m + kChunk - 1) / kChunk;
    // work item = (strip, row in strip, column chunk); strips that do not attain the maximum are skipped whole
    // (the strip maxima are tested 32 at a time: one dependent L2 load per strip and warp made this kernel 150 us)
    for (long long st0 = 0; st0 < nstrips; st0 += 32) {
      unsigned cand = __ballot_sync(0xffffffffu, st0 + lane < nstrips && strip_max[st0 + lane] == g);
      while (cand) {
        const long long st = st0 + (__ffs(cand) - 1);
        cand &= cand - 1;
        const long long items = (long long)kStripRows * nchunks;
        for (long long it = warp; it < items; it += nwarps) {
            const long long r = st * kStripRows + 1 + it / nchunks;
            if (r > n) continue;
            const long long c0 = 1 + (it % nchunks) * kChunk;
            const int32_t* Hr = H + r * pitch;
            for (long long j0 = c0; j0 < c0 + kChunk && j0 <= m; j0 += 32) {
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
    }
}

__global__ void finalize_kernel(const unsigned long long* key, const int* gmax, long long pitch, long long npairs,
                                long long* maxPos, int32_t* maxScore)
{
    const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= npairs) return;
    const int g = gmax[pair];
    long long pos = 0;
    if (g > 0) {
        const unsigned long long k = key[pair];
        const long long r = (long long)(0xffffffffu - (unsigned)(k & 0xffffffffu));
        const long long j = (long long)(k >> 32) - r;
        pos = r * pitch + j;
    }
    if (maxPos) maxPos[pair] = pos;
    if (maxScore) maxScore[pair] = g;
}

#ifndef SWB_FILL_ONLY
// ---------------------------------------------------------------------------------
// backtrack (omp_smithW.c:405-420): follow P from maxPos until a NONE cell, negating
// the path in place.  The chain is serial, so the kernel hides the HBM latency and keeps
// the per-cell dependency chain minimal (one shared-memory load + two ALU ops).
// The path moves up-left along a diagonal with small drift, so P is read in diagonal
// BANDS: 128 rows, and in row r the 128 columns around c0 - (i0 - r) where (i0, c0) is the
// cell the band was entered at.  While lane 0 of warp 0 walks the path inside the current
// band (in shared memory; it only records the visited band indices), warps 1..7 negate the
// cells of the previous band's path in global memory and prefetch the band that continues
// the diagonal with 16-byte cp.async (a band row is staged from the aligned global element
// at or before its first column; the walker tracks the row's 0..3 element shift off the
// dependency chain).  4-byte copies kept the LSU so busy that every load of the walker
// took ~190 clk.  The first and last column of a band and the row above it hold the marker
// 4: stepping on it ends the walk in this band.  If the walk leaves a band sideways, or
// enters the next one too far off centre, a band centred on the current cell is fetched
// instead.  One CTA of 256 threads, two band buffers (2 x 66 KB).
// ---------------------------------------------------------------------------------
constexpr int kBtRows = 128;
constexpr int kBtCols = 128;
constexpr int kBtRS = 132;                           // staged ints per band row: 33 aligned 16-byte chunks
constexpr int kBtThreads = 256;
constexpr int kBtFetchers = kBtThreads - 32;
constexpr int kBtBandInts = kBtRows * kBtRS;
constexpr int kBtPad = 160;                          // ints before each band = the marker row above it
constexpr int kBtList = 512;                         // a band holds at most 128 + 126 path cells
constexpr int kBtMark = 4;

// global index of band column 0 of band row rr, for the band entered at (i0, c0) (its cell (127, 64))
__device__ __forceinline__ long long bt_row_start(long long i0, long long c0, long long pitch, int rr, long long& r)
{
    r = i0 - (kBtRows - 1) + rr;
    return r * pitch + (c0 - (i0 - r) - kBtCols / 2);
}

// prefetch warps only.  limit = number of valid ints of P the kernel may read.
__device__ __forceinline__ void bt_fetch_band(int* buf, const int32_t* P, long long pitch, long long limit,
                                              long long i0, long long c0, unsigned long long* mbar, unsigned& phase)
{
    // One TMA bulk copy per band row (528 bytes from the 16-byte aligned global element at or before the row's first
    // column), issued by the first kBtRows fetcher threads and counted by an mbarrier.  The earlier version -- 33
    // cp.async of 16 bytes per row -- kept the SM's load/store pipeline so busy that every load of the walker took
    // about twice the shared-memory latency.
    const int tid = threadIdx.x - 32;
    const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
    if (tid < kBtRows) {
        const int rr = tid;
        long long r;
        const long long s = bt_row_start(i0, c0, pitch, rr, r);
        const long long g0 = s & ~3LL;                             // aligned global index of the row's first chunk
        int* dst = buf + rr * kBtRS;
        if (r >= 0 && g0 >= 0 && g0 + kBtRS <= limit) {
            const unsigned sd = (unsigned)__cvta_generic_to_shared(dst);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the buffer was last touched through the generic proxy
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(kBtRS * 4) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sd), "l"(P + g0), "r"(kBtRS * 4), "r"(mb) : "memory");
        } else {
            for (int k = 0; k < kBtRS; ++k) dst[k] = (r >= 0 && g0 + k >= 0 && g0 + k < limit) ? P[g0 + k] : 0;
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
        }
    }
    {
        unsigned done = 0;
        while (!done) {
            asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                         : "=r"(done) : "r"(mb), "r"(phase) : "memory");
        }
        phase ^= 1u;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kBtFetchers) : "memory");
    // A row landed with its first band column at offset (row start & 3): shift every row into place (one warp per
    // row, in place: all loads of the row precede its stores), so that band column k of band row rr is
    // buf[rr * kBtRS + k] and the walker moves with constant strides.
    {
        const int lane = tid & 31, fw = tid >> 5;
        for (int rr = fw; rr < kBtRows; rr += kBtFetchers / 32) {
            long long r;
            const int sh = (int)(bt_row_start(i0, c0, pitch, rr, r) & 3);
            int* row = buf + rr * kBtRS;
            const int v0 = row[sh + 4 * lane], v1 = row[sh + 4 * lane + 1], v2 = row[sh + 4 * lane + 2], v3 = row[sh + 4 * lane + 3];
            __syncwarp();
            // marker columns (band columns 0 and 127)
            *reinterpret_cast<int4*>(row + 4 * lane) = make_int4(lane == 0 ? kBtMark : v0, v1, v2, lane == 31 ? kBtMark : v3);
        }
    }
}

// negate the recorded path cells of a finished band in global memory (*= PATH, :417)
__device__ __forceinline__ void bt_writeback(const int* buf, const int* list, int count, int32_t* P, long long pitch,
                                             long long i0, long long c0)
{
    const int tid = threadIdx.x - 32;
    for (int e = tid; e < count; e += kBtFetchers) {
        const int a = list[e];
        const int rr = a / kBtRS;
        long long r;
        const long long s = bt_row_start(i0, c0, pitch, rr, r);
        P[s + (a - rr * kBtRS)] = -buf[a];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kBtFetchers) : "memory");   // all of it read before the buffer is reused
}

__global__ void __launch_bounds__(kBtThreads)
backtrack_kernel(int32_t* P, long long pitch, long long maxPos_arg,
                 const long long* d_maxPos, long long* d_pathLen, long long* d_endPos)
{
    extern __shared__ __align__(16) int bt_smem[];      // 2 x (pad + 128x132 ints), then 2 lists
    __shared__ long long s_len;
    __shared__ int s_done, s_a, s_count[2];
    __shared__ __align__(8) unsigned long long s_mbar[2];     // one per band buffer: counts the rows of a fetch
    unsigned mphase[2] = {0u, 0u};
    if (threadIdx.x == 0) {
        for (int k = 0; k < 2; ++k)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_mbar[k])), "r"(kBtRows) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long pos = d_maxPos ? *d_maxPos : maxPos_arg;
    if (pos <= 0) { if (threadIdx.x == 0) { if (d_pathLen) *d_pathLen = 0; if (d_endPos) *d_endPos = 0; } return; }
    long long i = pos / pitch, j = pos % pitch;         // current cell
    // the path only moves up and left: nothing after the end of the start row is ever needed (or read)
    const long long limit = (i + 1) * pitch;
    const bool walker = threadIdx.x == 0, fetcher = threadIdx.x >= 32;
    auto bandbuf = [&](int b) { return bt_smem + b * (kBtPad + kBtBandInts) + kBtPad; };
    auto listbuf = [&](int b) { return bt_smem + 2 * (kBtPad + kBtBandInts) + b * kBtList; };
    if (walker) { s_len = 0; s_done = 0; s_count[0] = 0; s_count[1] = 0; }
    for (int e = threadIdx.x; e < 2 * kBtPad; e += kBtThreads)
        bt_smem[(e / kBtPad) * (kBtPad + kBtBandInts) + e % kBtPad] = kBtMark;
    int cur = 0;
    long long bi = i, bc = j;                           // entry cell of the band in buffer `cur`
    long long pbi = 0, pbc = 0;                         // entry cell of the previous band (buffer cur^1) to write back
    bool have_prev = false;
#ifdef SWB_X_BTDEBUG
    long long dbg_walk = 0, dbg_total0 = clock64(); int dbg_bands = 0, dbg_miss = 0;
#endif
    if (fetcher) bt_fetch_band(bandbuf(cur), P, pitch, limit, bi, bc, &s_mbar[cur], mphase[cur]);
    __syncthreads();
    int k0 = kBtCols / 2;                               // band column of the current cell (bottom row)
    while (true) {
        // the band that continues this one's diagonal
        const long long ni = bi - kBtRows, nc = bc - kBtRows;
#ifdef SWB_X_BTSERIAL
        if (walker) {
#else
        if (fetcher) {
            if (have_prev) bt_writeback(bandbuf(cur ^ 1), listbuf(cur ^ 1), s_count[cur ^ 1], P, pitch, pbi, pbc);
            bt_fetch_band(bandbuf(cur ^ 1), P, pitch, limit, ni, nc, &s_mbar[cur ^ 1], mphase[cur ^ 1]);
        } else if (walker) {
#endif
#ifdef SWB_X_BTDEBUG
            const long long w0 = clock64(); ++dbg_bands;
#endif
            const int* band = bandbuf(cur);          // (not volatile: ptxas puts a YIELD into loops with volatile loads, ~150 clk per iteration)
            int* list = listbuf(cur);
            int a = (kBtRows - 1) * kBtRS + k0, cnt = 0;
            int pv = band[a];
            // step sizes in the staged band by P code (byte pv of the table): UP (:412) = one band row up and one
            // column right, LEFT (:414) = one column left, DIAGONAL (:410) = one band row up (the band follows the
            // diagonal); NONE, the marker and already negated cells select 0
            constexpr unsigned kTab = (unsigned)(kBtRS - 1) << 8 | 1u << 16 | (unsigned)kBtRS << 24;
            // (Tried: loading the eight cells reachable within two moves at once and resolving both moves with
            // selects -- 90 clk per cell instead of 68: a single thread pays several cycles per instruction, so the
            // walk is bound by its instruction count, not by the load.)
            while (true) {
                // speculative: address and load of the successor are issued before pv is checked, so the
                // branch resolves in the shadow of the load
                const int a2 = a - (int)__byte_perm(kTab, 0u, (unsigned)pv);
                const int pv2 = band[a2];
                if ((unsigned)(pv - 1) >= 3u) break;
                list[cnt++] = a;
                a = a2; pv = pv2;
            }
            s_count[cur] = cnt;
            s_len += cnt;
            s_done = (pv != kBtMark);                             // NONE (or an already negated cell) ends the path (:419)
            s_a = a;
#ifdef SWB_X_BTDEBUG
            dbg_walk += clock64() - w0;
#endif
        }
        __syncthreads();
#ifdef SWB_X_BTSERIAL
        if (fetcher) {
            if (have_prev) bt_writeback(bandbuf(cur ^ 1), listbuf(cur ^ 1), s_count[cur ^ 1], P, pitch, pbi, pbc);
            bt_fetch_band(bandbuf(cur ^ 1), P, pitch, limit, ni, nc, &s_mbar[cur ^ 1], mphase[cur ^ 1]);
        }
        __syncthreads();
#endif
        pbi = bi; pbc = bc; have_prev = true;
        {
            const int a = s_a;
            const int rr = (a + kBtRS) / kBtRS - 1;               // -1 = the marker row above the band
            long long r;
            const long long s = bt_row_start(bi, bc, pitch, rr, r);
            const long long g = s + (a - rr * kBtRS);
            i = r; j = g - r * pitch;
        }
        if (s_done) break;                                        // (i, j) = the cell that ended the walk
        // usable prefetch: the walk left through the top and enters the next band well inside it
        const long long kn = j - (nc - kBtCols / 2);              // its column in the next band's bottom row
        if (i == ni && kn >= 24 && kn <= kBtCols - 24) {
            cur ^= 1; bi = ni; bc = nc;
            k0 = (int)kn;
            __syncthreads();                                       // s_* read by everyone before the walker rewrites them
        } else {
#ifdef SWB_X_BTDEBUG
            ++dbg_miss;
#endif
            // the prefetched band is useless: write back the finished one now and fetch a band centred here
            __syncthreads();
            if (fetcher) {
                bt_writeback(bandbuf(cur), listbuf(cur), s_count[cur], P, pitch, pbi, pbc);
                bi = i; bc = j;
                bt_fetch_band(bandbuf(cur), P, pitch, limit, bi, bc, &s_mbar[cur], mphase[cur]);
            }
            have_prev = false;
            bi = i; bc = j;
            k0 = kBtCols / 2;
            __syncthreads();
        }
    }
    // the last band's path
    if (fetcher) bt_writeback(bandbuf(cur), listbuf(cur), s_count[cur], P, pitch, pbi, pbc);
    if (threadIdx.x == 0 && d_pathLen) *d_pathLen = s_len;
    if (threadIdx.x == 0 && d_endPos) *d_endPos = i * pitch + j;
#ifdef SWB_X_BTDEBUG
    if (threadIdx.x == 0) printf("backtrack: len %lld bands %d misses %d walk clk %lld total clk %lld\n", s_len, dbg_bands, dbg_miss, dbg_walk, clock64() - dbg_total0);
#endif
}

#endif  // SWB_FILL_ONLY

}  // namespace SWB_NS
