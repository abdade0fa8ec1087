// Developer microbenchmark: the fill kernel's STORE PATTERN without any compute -- 148 CTAs x 4 writer warps, each
// warp sweeps 32 rows of a 64-row strip left to right, one 128-byte line per row and matrix per round (row stride =
// pitch*4 bytes), strips taken from a ticket.  Answers: is the HBM write bandwidth reachable with this access pattern?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/ubench_pattern tools/ubench_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void pattern(int* H, int* P, long long pitch, long long n, long long m, int* ticket, int lines_per_visit)
{
    __shared__ int s_band;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_band = atomicAdd(ticket, 1);
        __syncthreads();
        const long long r0 = 1 + (long long)s_band * 128 + 32 * wid;      // 4 warps x 32 rows = 2 strips of 64 rows
        if (1 + (long long)s_band * 128 > n) return;
        const long long cols_per_visit = 32LL * lines_per_visit;
        // row pointers of my 32 rows (4 batches of 8), 128-byte aligned windows like the writers'
        for (long long c0 = 0; c0 + cols_per_visit <= m; c0 += cols_per_visit) {
#pragma unroll 1
            for (int b = 0; b < 32; b += 8) {
                int* hp[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    long long g = min(r0 + b + i, n) * pitch + c0;
                    hp[i] = H + (g - (g & 31)) + lane;
                }
                const long long pd = P - H;
                for (int k = 0; k < lines_per_visit; ++k) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { __stcs(hp[i] + 32 * k, (int)c0 + i); __stcs(hp[i] + 32 * k + pd, i & 3); }
                }
            }
        }
    }
}
int main()
{
    const long long m = 45000, n = 45000, pitch = m + 1;
    const size_t cells = (size_t)(n + 2) * pitch + 64;
    int *H, *P, *ticket;
    cudaMalloc(&H, cells * 4); cudaMalloc(&P, cells * 4); cudaMalloc(&ticket, 4);
    for (int lines : {1, 2, 4, 8}) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaMemset(ticket, 0, 4);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            pattern<<<148, 128>>>(H, P, pitch, n, m, ticket, lines);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("store pattern, %d x 128 B per row visit: %.3f ms = %.0f GB/s (%s)\n", lines, best, 8.0 * n * m / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    // more warps per SM (latency hiding for the stores): 148 x 8 CTAs of 128 threads
    for (int ctas : {2, 4, 8}) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaMemset(ticket, 0, 4);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            pattern<<<148 * ctas, 128>>>(H, P, pitch, n, m, ticket, 1);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("store pattern, 1 x 128 B per visit, %d CTAs per SM: %.3f ms = %.0f GB/s\n", ctas, best, 8.0 * n * m / best / 1e6);
    }
    return 0;
}
