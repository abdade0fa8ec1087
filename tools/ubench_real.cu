// Developer microbenchmark: the REAL Strip::step of swb_kernels.cuh, one warp, ring prefilled
// with valid tags, no writers: cost of the unrolled group body in isolation.
#include <cstdio>
#include "../smith-waterman_b200/csrc/swb_kernels.cuh"
using namespace swb;

// NOISE: 0 none, 1 = extra warps spin-polling shared memory (LDS.128 + nanosleep like a gate),
// 2 = extra warps doing writer-like work (LDS.32 from the staging ring + 2 coalesced global stores per row)
template <bool HASIN, bool OUT, int NOISE>
__global__ void k(const unsigned* aw_, int ngroups, long long* out_clk, int* sink, int* gbuf)
{
    extern __shared__ __align__(1024) int4 smem4[];
    __shared__ int4 ring[2 * kRing];
    __shared__ volatile int stop;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (threadIdx.x >= 32) {
        const int w = threadIdx.x >> 5;
        if (NOISE == 1) {
            const unsigned a = (unsigned)__cvta_generic_to_shared(ring) + 16u * 5;
            int4 v = lds_volatile_int4<0>(a);
            while (!stop) { __nanosleep(100); v = lds_volatile_int4<0>(a); if ((v.x & 3) == 3) break; }
        } else if (NOISE == 2) {
            const int* st = reinterpret_cast<const int*>(smem4);
            int r = 0, mx = 0;
            while (!stop) {
                for (int l = 0; l < 32; ++l) {
                    const int kk = st[l * kRowInts + ((32 * r + lane) & (kRowInts - 1))];
                    int* hp = gbuf + ((size_t)w * 64 + l) * 4096 + ((32 * r + lane) & 4095);
                    __stcs(hp, kk >> 4); __stcs(hp + 2048 * 4096, kk & 3); mx = max(mx, kk);
                }
                ++r;
            }
            if (mx == 0x7fffffff) *sink = mx;
        } else if (NOISE == 3) {                 // ALU only
            int x = w, y = lane;
            while (!stop) { for (int i = 0; i < 64; ++i) { x = x * 3 + y; y = (y ^ x) + 7; } }
            if (x == 0x7fffffff) *sink = x + y;
        } else if (NOISE == 4) {                 // shared-memory loads only (LDS.32, conflict-free)
            const int* st = reinterpret_cast<const int*>(smem4);
            int r = 0, mx = 0;
            while (!stop) { for (int l = 0; l < 32; ++l) mx = max(mx, st[l * kRowInts + ((32 * r + lane) & (kRowInts - 1))]); ++r; }
            if (mx == 0x7fffffff) *sink = mx;
        } else if (NOISE == 5) {                 // global stores only
            int r = 0;
            while (!stop) {
                for (int l = 0; l < 32; ++l) { int* hp = gbuf + ((size_t)w * 64 + l) * 4096 + ((32 * r + lane) & 4095); __stcs(hp, r); __stcs(hp + 2048 * 4096, l); }
                ++r;
            }
        } else if (NOISE == 6) {                 // shuffles only
            int x = lane;
            while (!stop) { for (int i = 0; i < 64; ++i) x = __shfl_up_sync(0xffffffffu, x, 1) + 1; }
            if (x == 0x7fffffff) *sink = x;
        }
        return;
    }
    for (int i = lane; i < 2 * kRing; i += 32) ring[i] = make_int4(0, 0, 0, 0);
    __syncwarp();
    Strip<64, true> S;
    constexpr int kT = 64, kRowInts = 256;
    S.lane = lane;
    for (int q = 0; q < kR; ++q) { S.b4[q] = 0x41414141u + 0x01010101u * ((lane + q) & 3); S.hl[q] = 0; }
    S.sm = opaque(16 * 3 + 7); S.sx = opaque(16 * -3 + 7); S.gu = opaque(16 * -2 + 5); S.gl = opaque(16 * -2 + 2);
    S.A0 = S.A1 = S.A2 = S.A3 = 0; S.dgp = 0;
    S.sa_base = (unsigned)__cvta_generic_to_shared(smem4 + (size_t)kR * lane * kT);
    S.sa = S.sa_base + 16u * lane;
    S.ring_in = (unsigned)__cvta_generic_to_shared(ring);
    S.ring_out = (unsigned)__cvta_generic_to_shared(ring + kRing);
    S.jmax = 1 << 30;
    S.has_in = opaque(HASIN ? 1 : 0);
    S.out_ring = opaque((OUT && lane == 31) ? 1 : 0);
    S.out_glob = opaque(0);
    S.gout = nullptr;
    const unsigned* aw = aw_ + kAPad - lane;
    unsigned cur[kGroup + 1], nxt[kGroup];
    for (int i = 0; i < kGroup; ++i) cur[i] = __ldg(aw + i);
    S.scores(cur[0]);
    long long total = 0;
    for (int g = 4; g < ngroups; ++g) {
        const int t0 = g * kGroup;
        for (int i = 0; i < kGroup; ++i) nxt[i] = __ldg(aw + t0 + kGroup + i);
        cur[kGroup] = nxt[0];
        const unsigned in_g  = S.ring_in + 16u * (unsigned)((t0 + 32) & (kRing - 1));
        const unsigned in_w  = S.ring_in + 16u * (unsigned)((t0 + 40) & (kRing - 1));
        const int want = 1 + (((t0 + 32) >> 6) & 1), want_w = 1 + (((t0 + 40) >> 6) & 1);
        const unsigned out_g = S.ring_out + 16u * (unsigned)((t0 & (kRing - 1)) + 1);
        const unsigned out_w = S.ring_out + 16u * (unsigned)((t0 + 8) & (kRing - 1));
        const int otag = 1 + ((t0 >> 6) & 1), otag_w = 1 + (((t0 + 8) >> 6) & 1);
        // make the 9 input entries of this group valid (what the producer would have done)
        if (HASIN) {
            if (lane < 9) {
                const int j = t0 + 1 + lane;
                ring[(j + 32) & (kRing - 1)] = make_int4(1 + (((j + 32) >> 6) & 1), 0, 0, 0);
            }
            __syncwarp();
        }
        const long long c0 = clock64();
#define SWB_STEP(E, I) S.template step<E, I>(t0 + I, cur[I + 1], in_g, in_w, want, want_w, out_g, out_w, otag, otag_w)
        SWB_STEP(false, 0); SWB_STEP(false, 1); SWB_STEP(false, 2); SWB_STEP(false, 3);
        SWB_STEP(false, 4); SWB_STEP(false, 5); SWB_STEP(false, 6); SWB_STEP(false, 7);
#undef SWB_STEP
        total += clock64() - c0;
        for (int i = 0; i < kGroup; ++i) cur[i] = nxt[i];
    }
    stop = 1;
    if (lane == 0) *out_clk = total;
    int acc = S.A0 + S.A1 + S.A2 + S.A3 + S.dgp + S.hl[0];
    if (acc == 0x7fffffff) *sink = acc;
}

template <bool HASIN, bool OUT, int NOISE>
void run(const unsigned* aw, int ngroups, long long* d_clk, int* d_sink, const char* name, int warps, int* gbuf)
{
    const size_t smem = (size_t)kR * 32 * kT * 16;
    cudaFuncSetAttribute(k<HASIN, OUT, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) k<HASIN, OUT, NOISE><<<1, 32 * warps, smem>>>(aw, ngroups, d_clk, d_sink, gbuf);
    long long clk = 0;
    cudaMemcpy(&clk, d_clk, sizeof clk, cudaMemcpyDeviceToHost);
    printf("kR=%d %-24s warps %2d %8.1f clk/step (%s)\n", kR, name, warps, (double)clk / ((ngroups - 4) * 8), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const int ngroups = 1024;
    unsigned* aw; long long* d_clk; int* d_sink;
    const int n = ngroups * 8 + 256;
    cudaMalloc(&aw, n * sizeof(unsigned)); cudaMalloc(&d_clk, 8); cudaMalloc(&d_sink, 4);
    unsigned* h = new unsigned[n];
    for (int i = 0; i < n; ++i) h[i] = 0x41414141u + 0x01010101u * (i & 3) + 0x00010000u * ((i >> 2) & 3);
    cudaMemcpy(aw, h, n * sizeof(unsigned), cudaMemcpyHostToDevice);
    int* gbuf; cudaMalloc(&gbuf, (size_t)4096 * 4096 * 4 * 4);
    run<false, false, 0>(aw, ngroups, d_clk, d_sink, "no in, no out", 1, gbuf);
    run<true, false, 0>(aw, ngroups, d_clk, d_sink, "in (valid), no out", 1, gbuf);
    run<true, true, 0>(aw, ngroups, d_clk, d_sink, "in (valid), out ring", 1, gbuf);
    for (int w : {2, 3, 5, 9, 13}) run<true, true, 1>(aw, ngroups, d_clk, d_sink, "in+out, pollers", w, gbuf);
    for (int w : {2, 3, 4, 5, 7}) run<true, true, 2>(aw, ngroups, d_clk, d_sink, "in+out, writer-like", w, gbuf);
    for (int w : {2, 4}) run<true, true, 3>(aw, ngroups, d_clk, d_sink, "in+out, ALU noise", w, gbuf);
    for (int w : {2, 4}) run<true, true, 4>(aw, ngroups, d_clk, d_sink, "in+out, LDS noise", w, gbuf);
    for (int w : {2, 4}) run<true, true, 5>(aw, ngroups, d_clk, d_sink, "in+out, STG noise", w, gbuf);
    for (int w : {2, 4}) run<true, true, 6>(aw, ngroups, d_clk, d_sink, "in+out, SHFL noise", w, gbuf);
    return 0;
}
