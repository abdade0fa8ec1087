// Developer microbenchmarks (not product code): the per-SM integer issue rates and the
// hand-off latencies that bound the wavefront kernel.  nvcc -arch=sm_100a -O3 -o tools/ubench.bin tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

// ---- throughput of independent VIADDMNMX / LOP3 / IMAD chains, per warp count
template <int OP>
__global__ void tput_kernel(int* out, long long* cyc, int seed)
{
    int v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = seed + k + threadIdx.x;
    const int c1 = seed | 1, c2 = seed * 3;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (OP == 0) v[k] = __viaddmax_s32(v[k], c1, c2);
            if (OP == 1) v[k] = (v[k] & c1) ^ c2;
            if (OP == 2) v[k] = v[k] * c1 + c2;
            if (OP == 3) { v[k] = __viaddmax_s32(v[k], c1, c2); v[k] = v[k] * c1 + c2; }   // ALU + FMA pipe mix
            if (OP == 4) v[k] = (v[k] > c2) ? c1 : v[k] + 1;                                 // ISETP+SEL-ish
            if (OP == 5) v[k] = v[k] >> 4;                                                   // SHF
        }
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- dependent latency: DPX chain, SHFL chain
__global__ void lat_kernel(int* out, long long* cyc, int seed)
{
    int v = seed + threadIdx.x;
    const int c1 = seed | 1, c2 = seed * 3;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v = __viaddmax_s32(v, c1, c2 + k);
    }
    long long t1 = clock64();
    int u = v;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) u = __shfl_up_sync(0xffffffffu, u, 1) + k;
    }
    long long t2 = clock64();
    int w = u;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) w = (__viaddmax_s32(w, c1, c2 + k)) & ~15;
    }
    long long t3 = clock64();
    out[threadIdx.x] = v + u + w;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
}

// ---- shared-memory ping-pong between two warps of one CTA (volatile flag, with/without fence)
template <bool FENCE>
__global__ void smem_pingpong(long long* cyc)
{
    __shared__ volatile int flag_a, flag_b;
    __shared__ int payload[64];
    if (threadIdx.x == 0) { flag_a = 0; flag_b = 0; }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int it = 1; it <= 1024; ++it) {
        if (w == 0) {
            payload[lane] = it;
            if (FENCE) __threadfence_block();
            if (lane == 0) flag_a = it;
            while (flag_b < it) { }
        } else {
            while (flag_a < it) { }
            if (FENCE) __threadfence_block();
            payload[32 + lane] = payload[lane];
            if (lane == 0) flag_b = it;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 1024;
}

// ---- global ping-pong between two CTAs (st.release.gpu / ld.acquire.gpu); cooperative launch not needed
// for 2 CTAs on an idle GPU, but guard with a timeout anyway.
__device__ __forceinline__ int ld_acq(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rel(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__global__ void gmem_pingpong(int* flags, int* payload, long long* cyc)
{
    const int lane = threadIdx.x;
    long long t0 = clock64();
    long long guard = 0;
    for (int it = 1; it <= 256; ++it) {
        if (blockIdx.x == 0) {
            payload[lane] = it;
            __syncwarp();
            if (lane == 0) st_rel(flags, it);
            while (ld_acq(flags + 32) < it) { if (++guard > (1LL << 24)) return; }
        } else {
            while (ld_acq(flags) < it) { if (++guard > (1LL << 24)) return; }
            payload[64 + lane] = payload[lane];
            __syncwarp();
            if (lane == 0) st_rel(flags + 32, it);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / 256;
}

// ---- write bandwidth: plain 16-byte streaming stores of the size of H+P at 45000^2
__global__ void write_kernel(int4* out, size_t n4)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int4 v = make_int4(1, 2, 3, 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) __stcs(out + i, v);
}

int main()
{
    int* out; long long* cyc; long long h[256];
    CK(cudaMalloc(&out, 1 << 22)); CK(cudaMalloc(&cyc, 256 * 8));
    const char* names[] = {"VIADDMNMX", "LOP3x2", "IMAD", "VIADDMNMX+IMAD", "ISETP+SEL", "SHF"};
    for (int warps : {1, 2, 4, 8, 16}) {
        for (int op = 0; op < 6; ++op) {
            switch (op) {
            case 0: tput_kernel<0><<<1, 32 * warps>>>(out, cyc, 3); break;
            case 1: tput_kernel<1><<<1, 32 * warps>>>(out, cyc, 3); break;
            case 2: tput_kernel<2><<<1, 32 * warps>>>(out, cyc, 3); break;
            case 3: tput_kernel<3><<<1, 32 * warps>>>(out, cyc, 3); break;
            case 4: tput_kernel<4><<<1, 32 * warps>>>(out, cyc, 3); break;
            case 5: tput_kernel<5><<<1, 32 * warps>>>(out, cyc, 3); break;
            }
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
            const double per = (double)h[0] / (ITERS * 8.0);
            printf("tput  %-16s warps/SM=%2d  %.2f clk per source-op per warp  => %.2f warp-ops/clk/SM\n", names[op], warps, per, warps / per);
        }
    }
    lat_kernel<<<1, 32>>>(out, cyc, 3);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost));
    printf("lat   VIADDMNMX dependent %.2f clk ; SHFL.UP+IADD dependent %.2f clk ; VIADDMNMX+LOP3 dependent %.2f clk\n",
           h[0] / (ITERS * 8.0), h[1] / (ITERS * 8.0), h[2] / (ITERS * 8.0));
    smem_pingpong<false><<<1, 64>>>(cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("smem  ping-pong round trip, no fence: %lld clk\n", h[0]);
    smem_pingpong<true><<<1, 64>>>(cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("smem  ping-pong round trip, __threadfence_block both sides: %lld clk\n", h[0]);
    int* flags; int* payload;
    CK(cudaMalloc(&flags, 1024)); CK(cudaMalloc(&payload, 1024)); CK(cudaMemset(flags, 0, 1024));
    gmem_pingpong<<<2, 32>>>(flags, payload, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
    printf("gmem  release/acquire ping-pong round trip between 2 CTAs: %lld clk\n", h[0]);

    // write bandwidth
    const size_t bytes = 16ull << 30;
    int4* big; CK(cudaMalloc(&big, bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        write_kernel<<<148 * 8, 512>>>(big, bytes / 16);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("write 16 GiB st.cs.v4: %.3f ms = %.1f GB/s\n", ms, bytes / (ms * 1e-3) / 1e9);
    }
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaMemsetAsync(big, 0, bytes));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("write 16 GiB cudaMemset: %.3f ms = %.1f GB/s\n", ms, bytes / (ms * 1e-3) / 1e9);
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device clock attr %d kHz\n", clk);
    return 0;
}
