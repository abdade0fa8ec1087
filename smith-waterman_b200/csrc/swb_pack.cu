// swb_pack.cu -- packed copy-back of H and P (include/swb200.h, "packed transfer").
//
// The reference keeps H and P as two int32 matrices in host memory (omp_smithW.c:203-216); a device fill therefore
// ends with 8 bytes per cell crossing PCIe, which takes 70 times longer than the fill itself (45000 x 45000: 16.2 GB,
// 313 ms at 51.8 GB/s against a 4.3 ms fill).  Both matrices are highly redundant on the wire:
//   * H is Lipschitz along a row: gap <= H[i][j] - H[i][j-1] <= match - gap (omp_smithW.c:331-388: the left move bounds
//     it below, dropping the last column of the best alignment bounds it above), i.e. -2 .. 5 for the default scoring;
//   * P holds a direction 0..3, negated on the path cells (omp_smithW.c:405-420): -3 .. 3.
// One byte per cell carries both: bits 7..3 = delta + 16 (-16 .. 15), bits 2..0 = P + 3.  The device packs rows (a pure
// HBM-bound pass: 8 bytes read, 1 written per cell), the bytes cross PCIe (1/8 of the traffic), and the host expands them
// into the caller's int32 matrices with a running sum per row, on several threads, chunk by chunk while later chunks are
// still in flight.  The int32 surface stays bit-exact; a matrix whose deltas do not fit (exotic scoring) raises a flag on
// the device and the caller falls back to the plain copies.
#include "../../include/swb200.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#if defined(__linux__)
#include <sched.h>
#endif
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

constexpr int kDeltaBias = 16, kPBias = 3;

// one block per row; thread t packs cells t, t+256, ...; four neighbouring lanes merge their bytes into one 32-bit store
__global__ void __launch_bounds__(256) pack_rows_kernel(const int32_t* __restrict__ H, const int32_t* __restrict__ P, long long pitch,
                                                        long long row0, long long cols, unsigned char* __restrict__ out,
                                                        long long out_pitch, int* __restrict__ overflow,
                                                        int32_t* __restrict__ row_base)
{
    const long long r = blockIdx.x;
    const int32_t* h = H + (row0 + r) * pitch;
    const int32_t* q = P + (row0 + r) * pitch;
    unsigned* o = reinterpret_cast<unsigned*>(out + r * out_pitch);
    const int lane = threadIdx.x & 31;
    bool bad = false;
    const long long jend = (cols + 31) & ~31ll;                  // whole warps take part in the shuffles
    for (long long j = threadIdx.x; j < jend; j += blockDim.x) {
        int hv = 0, pv = 0;
        if (j < cols) { hv = __ldcs(h + j); pv = __ldcs(q + j); }
        int prev = __shfl_up_sync(0xffffffffu, hv, 1);
        // column 0: its step counts from 0, or -- with a row base -- from itself (the base carries the value)
        if (lane == 0) prev = (j > 0 && j < cols) ? __ldg(h + j - 1) : (row_base != nullptr ? hv : 0);
        if (j == 0 && row_base != nullptr) row_base[r] = hv;
        const int d = hv - prev + kDeltaBias, pp = pv + kPBias;
        if (j < cols && (((unsigned)d > 31u) | ((unsigned)pp > 6u))) bad = true;
        unsigned w = j < cols ? ((((unsigned)d & 31u) << 3) | ((unsigned)pp & 7u)) << (8 * (lane & 3)) : 0u;
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        if ((lane & 3) == 0 && j < out_pitch) __stcs(o + (j >> 2), w);
    }
    if (bad) atomicOr(overflow, 1);
}

// ---- host side: one row of packed bytes -> int32 H and P
void expand_row_scalar(const unsigned char* src, long long cols, int32_t* h, int32_t* p, int acc)
{
    if (h && p) for (long long j = 0; j < cols; ++j) { const int v = src[j]; acc += (v >> 3) - kDeltaBias; h[j] = acc; p[j] = (v & 7) - kPBias; }
    else if (h) for (long long j = 0; j < cols; ++j) { acc += (src[j] >> 3) - kDeltaBias; h[j] = acc; }
    else if (p) for (long long j = 0; j < cols; ++j) p[j] = (src[j] & 7) - kPBias;
}

#if defined(__x86_64__)
// eight cells per iteration; non-temporal stores (the matrices are several times the last-level cache, a normal store
// would read every line before overwriting it)
__attribute__((target("avx2"))) void expand_row_avx2(const unsigned char* src, long long cols, int32_t* h, int32_t* p, int acc)
{
    long long j = 0;
    // scalar head until the output pointers are 32-byte aligned (H and P share the row offset; when their bases
    // differ mod 32 only one of them can be aligned: that one streams, the other uses unaligned stores)
    int32_t* lead = h ? h : p;
    while (j < cols && (reinterpret_cast<uintptr_t>(lead + j) & 31u) != 0) {
        const int v = src[j]; acc += (v >> 3) - kDeltaBias;
        if (h) h[j] = acc;
        if (p) p[j] = (v & 7) - kPBias;
        ++j;
    }
    const bool p_aligned = p && (reinterpret_cast<uintptr_t>(p + j) & 31u) == 0;
    __m256i carry = _mm256_set1_epi32(acc);
    const __m256i bias_d = _mm256_set1_epi32(kDeltaBias), bias_p = _mm256_set1_epi32(kPBias), m7 = _mm256_set1_epi32(7);
    const __m256i last = _mm256_set1_epi32(7);
    for (; j + 8 <= cols; j += 8) {
        const __m256i v = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(src + j)));
        if (h) {
            __m256i x = _mm256_sub_epi32(_mm256_srli_epi32(v, 3), bias_d);
            x = _mm256_add_epi32(x, _mm256_slli_si256(x, 4));
            x = _mm256_add_epi32(x, _mm256_slli_si256(x, 8));
            // low half's total into every element of the high half
            const __m256i lo_tot = _mm256_shuffle_epi32(_mm256_permute2x128_si256(x, x, 0x08), 0xff);
            x = _mm256_add_epi32(_mm256_add_epi32(x, lo_tot), carry);
            carry = _mm256_permutevar8x32_epi32(x, last);
            _mm256_stream_si256(reinterpret_cast<__m256i*>(h + j), x);
        }
        if (p) {
            const __m256i y = _mm256_sub_epi32(_mm256_and_si256(v, m7), bias_p);
            if (p_aligned || !h) _mm256_stream_si256(reinterpret_cast<__m256i*>(p + j), y);
            else _mm256_storeu_si256(reinterpret_cast<__m256i*>(p + j), y);
        }
    }
    acc = _mm256_extract_epi32(carry, 0);
    for (; j < cols; ++j) {
        const int v = src[j]; acc += (v >> 3) - kDeltaBias;
        if (h) h[j] = acc;
        if (p) p[j] = (v & 7) - kPBias;
    }
}
#endif

void expand_rows_range(const unsigned char* packed, long long packed_pitch, long long r_lo, long long r_hi, long long cols,
                       int32_t* H, int32_t* P, long long pitch, bool avx2, const int32_t* row_base)
{
    for (long long r = r_lo; r < r_hi; ++r) {
        const unsigned char* src = packed + r * packed_pitch;
        int32_t* h = H ? H + r * pitch : nullptr;
        int32_t* p = P ? P + r * pitch : nullptr;
        const int acc = row_base ? row_base[r] : 0;
#if defined(__x86_64__)
        if (avx2) { expand_row_avx2(src, cols, h, p, acc); continue; }
#endif
        expand_row_scalar(src, cols, h, p, acc);
    }
#if defined(__x86_64__)
    if (avx2) _mm_sfence();
#endif
}

bool have_avx2()
{
#if defined(__x86_64__)
    return __builtin_cpu_supports("avx2");
#else
    return false;
#endif
}

}  // namespace

extern "C" {

int swb_host_threads(void)
{
    if (const char* s = std::getenv("SWB_HOST_THREADS")) { const int v = std::atoi(s); if (v > 0) return std::min(v, 256); }
    unsigned hc = std::thread::hardware_concurrency();
#if defined(__linux__)
    // the cores this process may run on (a rank pinned to its GPU's NUMA node must not count the other node's)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) { const int c = CPU_COUNT(&set); if (c > 0) hc = (unsigned)c; }
#endif
    return (int)std::min(64u, std::max(1u, hc));
}

int64_t swb_packed_pitch(int64_t cols) { return (cols + 63) & ~(int64_t)63; }

int swb_pack_rows_async(const int32_t* dH, const int32_t* dP, int64_t pitch, int64_t row0, int64_t nrows, int64_t cols,
                        unsigned char* d_packed, int64_t packed_pitch, int* d_overflow, int32_t* d_row_base, int device,
                        void* stream)
{
    if (!dH || !dP || !d_packed || !d_overflow || nrows < 0 || cols <= 0 || pitch < cols || packed_pitch < cols ||
        (packed_pitch & 3) != 0 || row0 < 0)
        return SWB_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_packed) & 3u) return SWB_ERR_ALIGN;
    if (nrows == 0) return SWB_OK;
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) return SWB_ERR_CUDA;
    if (cur != device && cudaSetDevice(device) != cudaSuccess) return SWB_ERR_CUDA;
    int rc = SWB_OK;
    // (grid.x carries the rows: up to 2^31-1)
    pack_rows_kernel<<<(unsigned)nrows, 256, 0, static_cast<cudaStream_t>(stream)>>>(dH, dP, pitch, row0, cols, d_packed, packed_pitch,
                                                                                   d_overflow, d_row_base);
    if (cudaGetLastError() != cudaSuccess) rc = SWB_ERR_CUDA;
    if (cur != device) cudaSetDevice(cur);
    return rc;
}

int swb_expand_rows(const unsigned char* packed, int64_t packed_pitch, int64_t nrows, int64_t cols,
                    int32_t* H, int32_t* P, int64_t pitch, const int32_t* row_base, int threads)
{
    if (!packed || nrows < 0 || cols <= 0 || packed_pitch < cols || pitch < cols) return SWB_ERR_ARG;
    if (!H && !P) return SWB_OK;
    if (threads <= 0) threads = swb_host_threads();
    threads = (int)std::min<int64_t>(threads, std::max<int64_t>(1, nrows));
    const bool avx2 = have_avx2();
    if (threads == 1) { expand_rows_range(packed, packed_pitch, 0, nrows, cols, H, P, pitch, avx2, row_base); return SWB_OK; }
    std::vector<std::thread> pool;
    pool.reserve(threads);
    int started = 0;
    try {
        for (int t = 0; t < threads; ++t, ++started)
            pool.emplace_back([=] { expand_rows_range(packed, packed_pitch, nrows * t / threads, nrows * (t + 1) / threads, cols, H, P, pitch, avx2, row_base); });
    } catch (...) {
        // the host refused another thread: this thread does the slices that have no worker
        expand_rows_range(packed, packed_pitch, nrows * started / threads, nrows, cols, H, P, pitch, avx2, row_base);
    }
    for (auto& th : pool) th.join();
    return SWB_OK;
}

}  // extern "C"

// ---- pipelined form used by swb_ctx_align: chunk k is expanded by all threads as soon as its copy has landed
namespace swb_packed {

struct Chunk { long long r_lo, r_hi; cudaEvent_t ev; };

struct PackSource {                    // pack chunk by chunk in front of every copy (nullptr dH: the rows are packed already)
    const int32_t* dH = nullptr; const int32_t* dP = nullptr; long long pitch = 0;
    unsigned char* d_packed_w = nullptr; int* d_flag = nullptr; int32_t* d_base = nullptr; int32_t* h_base = nullptr;
};

// Copies the packed rows [0, nrows) from d_packed to h_packed in `nchunks` pieces on `st` (events recorded after each),
// and expands every piece into H / P on `threads` host threads while the later pieces are still in flight.  With a
// PackSource every piece is packed right in front of its copy, so the first piece is on its way after 1/nchunks of the
// packing pass instead of after all of it.
int copy_and_expand(const unsigned char* d_packed, unsigned char* h_packed, long long packed_pitch, long long nrows, long long cols,
                    int32_t* H, int32_t* P, long long pitch, const int32_t* row_base, cudaStream_t st, int device, int threads,
                    int nchunks, const PackSource* src = nullptr)
{
    if (threads <= 0) threads = swb_host_threads();
    nchunks = (int)std::max<long long>(1, std::min<long long>(nchunks, nrows));
    std::vector<Chunk> chunks(nchunks);
    cudaError_t err = cudaSuccess;
    for (int k = 0; k < nchunks && err == cudaSuccess; ++k) {
        Chunk& c = chunks[k];
        c.r_lo = nrows * k / nchunks; c.r_hi = nrows * (k + 1) / nchunks; c.ev = nullptr;
        err = cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming);
        if (err == cudaSuccess && src != nullptr && src->dH != nullptr) {
            if (swb_pack_rows_async(src->dH, src->dP, src->pitch, c.r_lo, c.r_hi - c.r_lo, cols, src->d_packed_w + c.r_lo * packed_pitch,
                                    packed_pitch, src->d_flag, src->d_base + c.r_lo, device, st) != SWB_OK)
                err = cudaErrorUnknown;
            if (err == cudaSuccess)
                err = cudaMemcpyAsync(src->h_base + c.r_lo, src->d_base + c.r_lo, (size_t)(c.r_hi - c.r_lo) * sizeof(int32_t),
                                      cudaMemcpyDeviceToHost, st);
        }
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(h_packed + c.r_lo * packed_pitch, d_packed + c.r_lo * packed_pitch,
                                  (size_t)(c.r_hi - c.r_lo) * packed_pitch, cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess) err = cudaEventRecord(c.ev, st);
    }
    std::atomic<int> failed{err == cudaSuccess ? 0 : 1};
    if (!failed.load()) {
        const bool avx2 = have_avx2();
        // worker t of `workers` expands slice t of every chunk as soon as the chunk has landed
        auto work = [&](int t_lo, int t_hi) {
            cudaSetDevice(device);
            for (int k = 0; k < nchunks; ++k) {
                if (cudaEventSynchronize(chunks[k].ev) != cudaSuccess) { failed.store(1); return; }
                const long long rows = chunks[k].r_hi - chunks[k].r_lo;
                expand_rows_range(h_packed, packed_pitch, chunks[k].r_lo + rows * t_lo / threads,
                                  chunks[k].r_lo + rows * t_hi / threads, cols, H, P, pitch, avx2, row_base);
            }
        };
        std::vector<std::thread> pool;
        pool.reserve(threads);
        int started = 0;
        try {
            for (int t = 0; t + 1 < threads; ++t, ++started) pool.emplace_back(work, t, t + 1);
        } catch (...) { }                                   // the host refused another thread: this one takes the rest
        work(started, threads);
        for (auto& th : pool) th.join();
    }
    for (auto& c : chunks) if (c.ev) cudaEventDestroy(c.ev);
    return failed.load() ? SWB_ERR_CUDA : SWB_OK;
}

}  // namespace swb_packed

// ---- whole delivery: rows [0, nrows) x columns [0, cols) of the device matrices into host int32 matrices
extern "C" {

size_t swb_d2h_packed_scratch_bytes(int64_t nrows, int64_t cols)
{
    if (nrows <= 0 || cols <= 0) return 0;
    const size_t packed = (size_t)swb_packed_pitch(cols) * (size_t)nrows;
    return packed + (((size_t)nrows * sizeof(int32_t) + 255) & ~(size_t)255) + 256;      // packed rows | row bases | flag
}

int swb_d2h_packed(const int32_t* dH, const int32_t* dP, int64_t pitch, int64_t nrows, int64_t cols,
                   int32_t* H, int32_t* P, int64_t host_pitch, void* d_scratch, void* h_scratch, int threads, int device,
                   void* stream)
{
    if (!dH || !dP || !d_scratch || !h_scratch || nrows <= 0 || cols <= 0 || pitch < cols || host_pitch < cols) return SWB_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_scratch) | reinterpret_cast<uintptr_t>(h_scratch)) & 15u) return SWB_ERR_ALIGN;
    if (!H && !P) return SWB_OK;
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) return SWB_ERR_CUDA;
    if (cur != device && cudaSetDevice(device) != cudaSuccess) return SWB_ERR_CUDA;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long pp = swb_packed_pitch(cols);
    const size_t packed_bytes = (size_t)pp * (size_t)nrows;
    const size_t base_bytes = ((size_t)nrows * sizeof(int32_t) + 255) & ~(size_t)255;
    unsigned char* d_packed = static_cast<unsigned char*>(d_scratch);
    unsigned char* h_packed = static_cast<unsigned char*>(h_scratch);
    int32_t* d_base = reinterpret_cast<int32_t*>(d_packed + packed_bytes);
    int32_t* h_base = reinterpret_cast<int32_t*>(h_packed + packed_bytes);
    int* d_flag = reinterpret_cast<int*>(d_packed + packed_bytes + base_bytes);
    int* h_flag = reinterpret_cast<int*>(h_packed + packed_bytes + base_bytes);
    int rc = SWB_OK;
    auto fail = [&](int code) { if (cur != device) cudaSetDevice(cur); return code; };
    if (cudaMemsetAsync(d_flag, 0, sizeof(int), st) != cudaSuccess) return fail(SWB_ERR_CUDA);
    // pack, copy and expand chunk by chunk; whether every value fitted the format is known only after the last
    // chunk -- if one did not (exotic scoring), the plain copies below overwrite whatever was expanded
    swb_packed::PackSource src;
    src.dH = dH; src.dP = dP; src.pitch = pitch; src.d_packed_w = d_packed; src.d_flag = d_flag; src.d_base = d_base; src.h_base = h_base;
    rc = swb_packed::copy_and_expand(d_packed, h_packed, pp, nrows, cols, H, P, host_pitch, h_base, st, device, threads, 24, &src);
    if (rc != SWB_OK) return fail(rc);
    if (cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return fail(SWB_ERR_CUDA);
    if (*h_flag != 0) {
        cudaError_t e = cudaSuccess;
        if (H) e = cudaMemcpy2DAsync(H, (size_t)host_pitch * 4, dH, (size_t)pitch * 4, (size_t)cols * 4, (size_t)nrows, cudaMemcpyDeviceToHost, st);
        if (P && e == cudaSuccess) e = cudaMemcpy2DAsync(P, (size_t)host_pitch * 4, dP, (size_t)pitch * 4, (size_t)cols * 4, (size_t)nrows, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = SWB_ERR_CUDA;
    }
    return fail(rc);
}

}  // extern "C"
