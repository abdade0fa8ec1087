// Developer microbenchmark: which pipe does IDP.4A (dp4a) share, and what are its rate and latency?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/ubench_dp4a tools/ubench_dp4a.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int MIX>
__global__ void k(int seed, long long* clk, int* sink)
{
    const int one = __shfl_sync(0xffffffffu, 1, 0);
    int v[8], w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = seed + i + threadIdx.x; w[i] = seed * 3 + i; }
    const int c1 = seed | 1, c2 = seed * 5;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MIX == 0) v[i] = __dp4a(v[i], 0x00001000, c2);                                   // IDP alone (independent chains)
            if (MIX == 1) { v[i] = __dp4a(v[i], 0x00001000, c2); w[i] = __viaddmax_s32(w[i], c1, c2); }   // IDP + ALU
            if (MIX == 2) { v[i] = __dp4a(v[i], 0x00001000, c2); asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(w[i]) : "r"(w[i]), "r"(one), "r"(c1)); }  // IDP + IMAD
            if (MIX == 3) w[i] = __viaddmax_s32(w[i], c1, c2);                                   // ALU alone
            if (MIX == 4) asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(w[i]) : "r"(w[i]), "r"(one), "r"(c1));    // IMAD alone
            if (MIX == 5) { w[i] = __viaddmax_s32(w[i], c1, c2); asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(v[i]) : "r"(v[i]), "r"(one), "r"(c1)); }  // ALU + IMAD
            if (MIX == 6) v[i] = __byte_perm(v[i], c1, c2 + i);                                  // PRMT alone
            if (MIX == 7) { v[i] = __byte_perm(v[i], c1, c2 + i); w[i] = __viaddmax_s32(w[i], c1, c2); }   // PRMT + ALU
        }
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i] + w[i];
    if (threadIdx.x == 0) *clk = t1 - t0;
    if (s == 0x7fffffff) *sink = s;
}
__global__ void lat(int seed, long long* clk, int* sink)
{
    int v = seed + threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v = __dp4a(v, 0x00000110, seed + i);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) *clk = t1 - t0;
    if (v == 0x7fffffff) *sink = v;
}
template <int MIX> void run(const char* name, int ops, long long* d, int* s)
{
    for (int r = 0; r < 2; ++r) k<MIX><<<1, 32>>>(12345, d, s);
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %6.2f clk per loop body instruction pair/op (1 warp, 8 independent chains; %d op(s) per body) -> %.2f clk per instruction\n",
           name, (double)c / (ITERS * 8.0), ops, (double)c / (ITERS * 8.0 * ops));
}
int main()
{
    long long* d; int* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
    run<0>("IDP.4A alone", 1, d, s); run<3>("VIADDMNMX alone", 1, d, s); run<4>("IMAD alone", 1, d, s); run<6>("PRMT alone", 1, d, s);
    run<1>("IDP.4A + VIADDMNMX", 2, d, s); run<2>("IDP.4A + IMAD", 2, d, s); run<5>("VIADDMNMX + IMAD", 2, d, s); run<7>("PRMT + VIADDMNMX", 2, d, s);
    for (int r = 0; r < 2; ++r) lat<<<1, 32>>>(12345, d, s);
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("IDP.4A dependent latency     %6.2f clk\n", (double)c / (ITERS * 8.0));
    return 0;
}
