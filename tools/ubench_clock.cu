// Developer tool: SM clock actually delivered during short kernels (clock64 vs globaltimer).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(long long cycles, unsigned long long* out)
{
    unsigned long long g0, g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long c0 = clock64();
    while (clock64() - c0 < cycles) { }
    const long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = c1 - c0; out[1] = g1 - g0; }
}
int main()
{
    unsigned long long* d; cudaMalloc(&d, 16);
    for (int blocks : {1, 148, 1184}) for (long long cyc : {200000LL, 2000000LL, 20000000LL, 200000000LL}) {
        for (int rep = 0; rep < 3; ++rep) {
            spin<<<blocks, 128>>>(cyc, d);
            unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("blocks %4d cycles %10lld rep %d: %.0f MHz (%.3f ms)\n", blocks, cyc, rep, 1e3 * h[0] / (double)h[1], h[1] * 1e-6);
        }
    }
    return 0;
}
