// Developer microbenchmark: cost of one compute step of the fill kernel in isolation
// (one warp per SM, no polling, no writers).  Variants remove one ingredient at a time.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/ubench_step tools/ubench_step.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kNone = 8;

template <int R, int VAR>
__global__ void step_kernel(const unsigned* __restrict__ aw, int nsteps, int sm_, int sx_, int gu_, int gl_,
                            long long* out_clk, int* sink)
{
    extern __shared__ int4 stage[];
    const int lane = threadIdx.x & 31;
    int sm = __shfl_sync(0xffffffffu, sm_, 0), sx = __shfl_sync(0xffffffffu, sx_, 0);
    int gu = __shfl_sync(0xffffffffu, gu_, 0), gl = __shfl_sync(0xffffffffu, gl_, 0);
    unsigned b4[R];
    int hl[R];
#pragma unroll
    for (int q = 0; q < R; ++q) { b4[q] = 0x41414141u + 0x01010101u * ((lane + q) & 3); hl[q] = 0; }
    int A0 = 0, A1 = 0, A2 = 0, A3 = 0, dgp = 0;
    int s[R][4];
#pragma unroll
    for (int q = 0; q < R; ++q) { s[q][0] = s[q][1] = s[q][2] = s[q][3] = sx; }
    unsigned sa_base = (unsigned)__cvta_generic_to_shared(stage + (size_t)R * lane * 64);
    unsigned sa = sa_base + 16u * lane;
    __shared__ int4 ring[64];
    __shared__ int flag;
    for (int i = lane; i < 64; i += 32) ring[i] = make_int4(1, 0, 0, 0);     // always-valid tag 1
    __syncwarp();
    const unsigned ring_a = (unsigned)__cvta_generic_to_shared(ring);
    const int out_on = (lane == 31) ? 1 : 0;
    const unsigned* p = aw + 64 - lane;
    unsigned word = __ldg(p);
    const long long c0 = clock64();
#pragma unroll 8
    for (int t = 0; t < nsteps; ++t) {
        const unsigned next_word = __ldg(p + t + 1);
        int4 v = make_int4(0, 0, 0, 0);
        if (VAR >= 4) asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ring_a + 16u * ((t + 1) & 63)) : "memory");
        int u0 = A0, u1 = A1, u2 = A2, u3 = A3, dg = dgp;
        dgp = A3;
        int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int p0 = __viaddmax_s32(dg, s[q][0], kNone);
            const int p1 = __viaddmax_s32(u0, s[q][1], kNone);
            const int p2 = __viaddmax_s32(u1, s[q][2], kNone);
            const int p3 = __viaddmax_s32(u2, s[q][3], kNone);
            const int t0 = __viaddmax_s32(u0, gu, p0);
            const int t1 = __viaddmax_s32(u1, gu, p1);
            const int t2 = __viaddmax_s32(u2, gu, p2);
            const int t3 = __viaddmax_s32(u3, gu, p3);
            dg = hl[q];
            const int k0 = __viaddmax_s32(hl[q], gl, t0);
            const int h0 = k0 & ~15;
            if (VAR != 3 && q == R - 1) n0 = __shfl_up_sync(0xffffffffu, h0, 1);
            const int k1 = __viaddmax_s32(h0, gl, t1);
            const int h1 = k1 & ~15;
            if (VAR != 3 && q == R - 1) n1 = __shfl_up_sync(0xffffffffu, h1, 1);
            const int k2 = __viaddmax_s32(h1, gl, t2);
            const int h2 = k2 & ~15;
            if (VAR != 3 && q == R - 1) n2 = __shfl_up_sync(0xffffffffu, h2, 1);
            const int k3 = __viaddmax_s32(h2, gl, t3);
            const int h3 = k3 & ~15;
            if (VAR != 3 && q == R - 1) n3 = __shfl_up_sync(0xffffffffu, h3, 1);
            hl[q] = h3;
            if (VAR != 1)
                asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(sa + 1024u * q), "r"(k0), "r"(k1), "r"(k2), "r"(k3) : "memory");
            u0 = h0; u1 = h1; u2 = h2; u3 = h3;
        }
        sa = ((sa + 16u) & 1023u) | sa_base;
        if (VAR != 2) {
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const unsigned x = next_word ^ b4[q];
                s[q][0] = (x & 0x000000ffu) ? sx : sm;
                s[q][1] = (x & 0x0000ff00u) ? sx : sm;
                s[q][2] = (x & 0x00ff0000u) ? sx : sm;
                s[q][3] = (x & 0xff000000u) ? sx : sm;
            }
        }
        if (VAR == 3) { n0 = u0; n1 = u1; n2 = u2; n3 = u3; }
        if (VAR >= 5) {
            asm volatile("{ .reg .pred q; setp.ne.s32 q, %5, 0; @q st.volatile.shared.v4.s32 [%0], {%1,%2,%3,%4}; }"
                         ::"r"(ring_a + 16u * (t & 63)), "r"(u0 | 1), "r"(u1), "r"(u2), "r"(u3), "r"(out_on) : "memory");
        }
        if (VAR >= 4) {
            if (__builtin_expect((v.x & 3) != 1, 0)) {
                do { asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ring_a + 16u * ((t + 1) & 63)) : "memory"); } while ((v.x & 3) != 1);
            }
        }
        if (VAR >= 6 && (t & 7) == 7) { __syncwarp(); __threadfence_block(); if (lane == 0) *(volatile int*)&flag = t; }
        if (false) { __syncwarp(); if (lane == 0) *(volatile int*)&flag = t; }
        const bool l0 = lane == 0;
        A0 = l0 ? (v.x & ~15) : n0; A1 = l0 ? v.y : n1; A2 = l0 ? v.z : n2; A3 = l0 ? v.w : n3;
        word = next_word;
    }
    const long long c1 = clock64();
    if (lane == 0 && blockIdx.x == 0) *out_clk = c1 - c0;
    int acc = A0 + A1 + A2 + A3 + dgp + (int)word;
#pragma unroll
    for (int q = 0; q < R; ++q) acc += hl[q] + s[q][0];
    if (acc == 0x7fffffff) *sink = acc;
}

template <int R, int VAR>
void run(const unsigned* aw, int nsteps, long long* d_clk, int* d_sink, const char* name)
{
    const size_t smem = (size_t)R * 32 * 1024;
    cudaFuncSetAttribute(step_kernel<R, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) step_kernel<R, VAR><<<1, 32, smem>>>(aw, nsteps, 16 * 3 + 7, 16 * -3 + 7, 16 * -2 + 5, 16 * -2 + 2, d_clk, d_sink);
    long long clk = 0;
    cudaMemcpy(&clk, d_clk, sizeof clk, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("R=%d %-28s %8.1f clk/step  %6.1f clk/cell-column  (%s)\n", R, name, (double)clk / nsteps, (double)clk / nsteps / (4 * R),
           cudaGetErrorString(e));
}

int main()
{
    const int nsteps = 8192;
    unsigned* aw; long long* d_clk; int* d_sink;
    cudaMalloc(&aw, (nsteps + 256) * sizeof(unsigned));
    cudaMalloc(&d_clk, 8); cudaMalloc(&d_sink, 4);
    unsigned* h = new unsigned[nsteps + 256];
    for (int i = 0; i < nsteps + 256; ++i) h[i] = 0x41414141u + 0x01010101u * (i & 3) + 0x00010000u * ((i >> 2) & 3);
    cudaMemcpy(aw, h, (nsteps + 256) * sizeof(unsigned), cudaMemcpyHostToDevice);
#define ALL(R) run<R, 0>(aw, nsteps, d_clk, d_sink, "full step"); run<R, 1>(aw, nsteps, d_clk, d_sink, "no staging store"); \
               run<R, 2>(aw, nsteps, d_clk, d_sink, "no score selects"); run<R, 3>(aw, nsteps, d_clk, d_sink, "no shuffles"); \
               run<R, 4>(aw, nsteps, d_clk, d_sink, "+ ring poll (valid)"); run<R, 5>(aw, nsteps, d_clk, d_sink, "+ poll + out store"); \
               run<R, 6>(aw, nsteps, d_clk, d_sink, "+ poll + out + group fence");
    ALL(1) ALL(2) ALL(4)
    return 0;
}
