// Developer microbenchmark (round 2): the profile-based compute step in isolation (one warp per SM, no polling, no
// writers), variants that remove one ingredient at a time, and the dependent-issue latencies of the instructions on
// the step's recurrence.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/ubench_step2 tools/ubench_step2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kNone = 8;

__device__ __forceinline__ int addf(int x, int one, int y)
{
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(y));
    return d;
}
__device__ __forceinline__ int4 lds4(unsigned a)
{
    int4 v;
    asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

// VAR bits: 1 no shuffles, 2 no staging store, 4 no profile loads, 8 no ring load / lane-0 select, 16 no hand-off store
template <int R, int VAR, int MEM = 0>
__global__ void step_kernel(int nsteps, int gu_, int gl_, long long* out_clk, int* sink)
{
    extern __shared__ int4 smem[];
    int4* stage = smem;                       // [R*32 rows][64]
    int4* prof = smem + R * 32 * 64;          // [4][272]
    __shared__ int4 ring[64];
    const int lane = threadIdx.x & 31;
    const int one = __shfl_sync(0xffffffffu, 1, 0);
    const int gu = __shfl_sync(0xffffffffu, gu_, 0), gl = __shfl_sync(0xffffffffu, gl_, 0);
    for (int i = lane; i < 4 * 272; i += 32) { const int m = ((i * 7 + (i >> 3)) & 3) == (i / 272) ? 55 : -41; prof[i] = make_int4(m, -41, m, -41); }
    for (int i = lane; i < 64; i += 32) ring[i] = make_int4(1, 0, 0, 0);
    __syncwarp();
    int hl[R];
#pragma unroll
    for (int q = 0; q < R; ++q) hl[q] = 0;
    int A0 = 0, A1 = 0, A2 = 0, A3 = 0, dgp = 0;
    unsigned rb[R];
#pragma unroll
    for (int q = 0; q < R; ++q) rb[q] = (unsigned)__cvta_generic_to_shared(prof + ((lane + q) & 3) * 272);
    const unsigned ring_a = (unsigned)__cvta_generic_to_shared(ring);
    unsigned sa_base = (unsigned)__cvta_generic_to_shared(stage + (size_t)R * lane * 64);
    unsigned sa = sa_base + 16u * lane;
    const int out_on = (lane == 31) ? 1 : 0;
    int4 pf[R][2];
#pragma unroll
    for (int q = 0; q < R; ++q) { pf[q][0] = lds4(rb[q]); pf[q][1] = lds4(rb[q] + 16); }
    const long long c0 = clock64();
#pragma unroll 8
    for (int t = 0; t < nsteps; ++t) {
        int4 v = make_int4(0, 0, 0, 0);
        if (!(VAR & 8)) v = MEM ? ring[(t + 1) & 63] : lds4(ring_a + 16u * ((t + 1) & 63));
        int4 sc[R];
        int4 kk[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            sc[q] = pf[q][t & 1];
            if (!(VAR & 4)) pf[q][t & 1] = MEM ? prof[((lane + q) & 3) * 272 + ((t + 2) & 255)] : lds4(rb[q] + 16u * ((t + 2) & 255));
        }
        int u0 = A0, u1 = A1, u2 = A2, u3 = A3, dg = dgp;
        dgp = A3;
        int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int d0 = addf(dg, one, sc[q].x), d1 = addf(u0, one, sc[q].y), d2 = addf(u1, one, sc[q].z), d3 = addf(u2, one, sc[q].w);
            const int v0 = addf(u0, one, gu), v1 = addf(u1, one, gu), v2 = addf(u2, one, gu), v3 = addf(u3, one, gu);
            const int t0 = __vimax3_s32(d0, v0, kNone), t1 = __vimax3_s32(d1, v1, kNone), t2 = __vimax3_s32(d2, v2, kNone), t3 = __vimax3_s32(d3, v3, kNone);
            dg = hl[q];
            const int k0 = __viaddmax_s32(hl[q], gl, t0); const int h0 = k0 & ~15;
            if (!(VAR & 1) && q == R - 1) n0 = __shfl_up_sync(0xffffffffu, h0, 1);
            const int k1 = __viaddmax_s32(h0, gl, t1); const int h1 = k1 & ~15;
            if (!(VAR & 1) && q == R - 1) n1 = __shfl_up_sync(0xffffffffu, h1, 1);
            const int k2 = __viaddmax_s32(h1, gl, t2); const int h2 = k2 & ~15;
            if (!(VAR & 1) && q == R - 1) n2 = __shfl_up_sync(0xffffffffu, h2, 1);
            const int k3 = __viaddmax_s32(h2, gl, t3); const int h3 = k3 & ~15;
            if (!(VAR & 1) && q == R - 1) n3 = __shfl_up_sync(0xffffffffu, h3, 1);
            hl[q] = h3;
            kk[q] = make_int4(k0, k1, k2, k3);
            if (!(VAR & 2) && MEM == 0)
                asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(sa + 1024u * q), "r"(k0), "r"(k1), "r"(k2), "r"(k3) : "memory");
            if (!(VAR & 2) && MEM == 1) stage[((size_t)R * lane + q) * 64 + ((t + lane) & 63)] = kk[q];
            u0 = h0; u1 = h1; u2 = h2; u3 = h3;
        }
        sa = ((sa + 16u) & 1023u) | sa_base;
        if (VAR & 1) { n0 = u0; n1 = u1; n2 = u2; n3 = u3; }
        if (!(VAR & 2) && MEM == 2) {
#pragma unroll
            for (int q = 0; q < R; ++q) stage[((size_t)R * lane + q) * 64 + ((t + lane) & 63)] = kk[q];
        }
        if (!(VAR & 16) && MEM) { if (out_on) ring[t & 63] = make_int4(u0 | 1, u1, u2, u3); }
        if (!(VAR & 16) && !MEM)
            asm volatile("{ .reg .pred q; setp.ne.s32 q, %5, 0; @q st.volatile.shared.v4.s32 [%0], {%1,%2,%3,%4}; }"
                         ::"r"(ring_a + 16u * (t & 63)), "r"(u0 | 1), "r"(u1), "r"(u2), "r"(u3), "r"(out_on) : "memory");
        const bool l0 = lane == 0;
        if (VAR & 8) { A0 = n0; A1 = n1; A2 = n2; A3 = n3; }
        else { A0 = l0 ? (v.x & ~15) : n0; A1 = l0 ? v.y : n1; A2 = l0 ? v.z : n2; A3 = l0 ? v.w : n3; }
    }
    const long long c1 = clock64();
    if (lane == 0 && blockIdx.x == 0) *out_clk = c1 - c0;
    int acc = A0 + A1 + A2 + A3 + dgp;
#pragma unroll
    for (int q = 0; q < R; ++q) acc += hl[q] + pf[q][0].x + pf[q][1].y;
    if (acc == 0x7fffffff) *sink = acc;
}

// dependent-issue latency of one instruction kind (OP) in a chain of 8 per iteration
template <int OP>
__global__ void lat_kernel(int seed, long long* out_clk, int* sink)
{
    const int one = __shfl_sync(0xffffffffu, 1, 0);
    int v = seed + threadIdx.x, w = seed * 3;
    const int c1 = seed | 1, c2 = seed * 5;
    const long long t0 = clock64();
    for (int it = 0; it < 2048; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (OP == 0) v = __viaddmax_s32(v, c1, c2 + k);
            if (OP == 1) v = __vimax3_s32(v, c1 + k, c2);
            if (OP == 2) v = addf(v, one, c1 + k);
            if (OP == 3) v = (v & ~15) ^ k;
            if (OP == 4) v = __shfl_up_sync(0xffffffffu, v, 1);
            if (OP == 5) v = (threadIdx.x == 0) ? w : v + k;                      // SEL-ish
            if (OP == 6) { v = addf(v, one, c1); v = __vimax3_s32(v, c2, kNone); v = __viaddmax_s32(w, c1, v); v &= ~15; }   // one cell on the vertical chain
            if (OP == 7) { v = __shfl_up_sync(0xffffffffu, v, 1); v = (threadIdx.x == 0) ? w : v; }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) *out_clk = t1 - t0;
    if (v == 0x7fffffff) *sink = v;
}

template <int R, int VAR, int MEM = 0>
void run(int nsteps, long long* d_clk, int* d_sink, const char* name)
{
    const size_t smem = (size_t)R * 32 * 1024 + 4 * 272 * 16;
    cudaFuncSetAttribute(step_kernel<R, VAR, MEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) step_kernel<R, VAR, MEM><<<1, 32, smem>>>(nsteps, 16 * -2 + 5, 16 * -2 + 2, d_clk, d_sink);
    long long clk = 0;
    cudaMemcpy(&clk, d_clk, sizeof clk, cudaMemcpyDeviceToHost);
    printf("R=%d %-44s %8.1f clk/step  %6.2f clk/cell   (%s)\n", R, name, (double)clk / nsteps, (double)clk / nsteps / (4 * R),
           cudaGetErrorString(cudaGetLastError()));
}
template <int OP>
void lat(long long* d_clk, int* d_sink, const char* name, int per)
{
    for (int rep = 0; rep < 2; ++rep) lat_kernel<OP><<<1, 32>>>(12345, d_clk, d_sink);
    long long clk = 0;
    cudaMemcpy(&clk, d_clk, sizeof clk, cudaMemcpyDeviceToHost);
    printf("lat  %-52s %6.2f clk per link (%s)\n", name, (double)clk / (2048.0 * 8 * per), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const int nsteps = 8192;
    long long* d_clk; int* d_sink;
    cudaMalloc(&d_clk, 8); cudaMalloc(&d_sink, 4);
    lat<0>(d_clk, d_sink, "VIADDMNMX", 1); lat<1>(d_clk, d_sink, "VIMNMX3", 1); lat<2>(d_clk, d_sink, "IMAD (register multiplier)", 1);
    lat<3>(d_clk, d_sink, "LOP3 (and, xor)", 1); lat<4>(d_clk, d_sink, "SHFL.UP", 1); lat<5>(d_clk, d_sink, "ISETP/SEL + IADD", 1);
    lat<6>(d_clk, d_sink, "cell: IMAD -> VIMNMX3 -> VIADDMNMX -> LOP3", 1); lat<7>(d_clk, d_sink, "SHFL.UP -> SEL", 1);
#define ALL(R) run<R, 0>(nsteps, d_clk, d_sink, "full profile step"); run<R, 1>(nsteps, d_clk, d_sink, "no shuffles"); \
               run<R, 2>(nsteps, d_clk, d_sink, "no staging store"); run<R, 4>(nsteps, d_clk, d_sink, "no profile loads"); \
               run<R, 8>(nsteps, d_clk, d_sink, "no ring load / lane-0 select"); run<R, 16>(nsteps, d_clk, d_sink, "no hand-off store"); \
               run<R, 30>(nsteps, d_clk, d_sink, "cells + shuffles only"); run<R, 31>(nsteps, d_clk, d_sink, "cells only");
    ALL(2)
    run<2, 0, 1>(nsteps, d_clk, d_sink, "full step, plain C++ shared accesses");
    run<2, 0, 2>(nsteps, d_clk, d_sink, "full step, plain + staging at end of step");
    run<2, 4, 1>(nsteps, d_clk, d_sink, "plain, no profile loads");
    run<2, 2, 1>(nsteps, d_clk, d_sink, "plain, no staging");
    run<2, 1, 1>(nsteps, d_clk, d_sink, "plain, no shuffles");
    run<1, 0, 1>(nsteps, d_clk, d_sink, "R=1 full step, plain");
    run<4, 0, 1>(nsteps, d_clk, d_sink, "R=4 full step, plain");
    run<3, 0, 0>(nsteps, d_clk, d_sink, "R=3 full step");
    run<3, 0, 1>(nsteps, d_clk, d_sink, "R=3 full step, plain");
    return 0;
}
