// swb_api.cu -- thin C ABI over the sm_100a kernels (include/swb200.h).
// Host code is plain C++/CUDA runtime; no torch types, no CPU fallback.
#include "../../include/swb200.h"
#define SWB_MERGED_FORMS 1              // (the batch geometry launches ONE fill kernel that holds both forms of the cell arithmetic)
#include "swb_kernels.cuh"              // namespace swb: two rows per lane (batches, score-only) + backtrack, argmax ...
#undef SWB_MERGED_FORMS
#define SWB_MERGED_FORMS 0
#include "swb_backtrack.cuh"
#undef SWB_NS
#undef SWB_ROWS_PER_LANE
#define SWB_NS swb_tall
#define SWB_ROWS_PER_LANE 3
#define SWB_FILL_ONLY 1
#ifndef SWB_TALL_HALF_SKEW
#define SWB_TALL_HALF_SKEW 1
#endif
#undef SWB_HALF_SKEW
#define SWB_HALF_SKEW SWB_TALL_HALF_SKEW
#include "swb_kernels.cuh"              // namespace swb_tall: three rows per lane (single large pairs)
#undef SWB_NS
#undef SWB_HALF_SKEW
#define SWB_NS swb_wide
#define SWB_HALF_SKEW 0
#include "swb_kernels.cuh"              // namespace swb_wide: three rows per lane, full skew (single pairs far wider than tall)
#undef SWB_NS
#undef SWB_FILL_ONLY
#undef SWB_HALF_SKEW

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#ifndef SWB_BT2
#define SWB_BT2 0                      // 1: every backtrack runs the jump-table kernel (swb_backtrack.cuh); 0: only the ones
#endif                                 //    that emit moves do (measured equal: 2.0 ms for 53 856 cells)

namespace {

thread_local char g_cuda_err[512] = "";

int cuda_fail(cudaError_t e, const char* what, int line)
{
    std::snprintf(g_cuda_err, sizeof g_cuda_err, "%s failed at swb_api.cu:%d: %s (%s)", what, line,
                  cudaGetErrorName(e), cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? SWB_ERR_NOMEM : SWB_ERR_CUDA;
}

#define SWB_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call, __LINE__);         \
    } while (0)

struct DeviceGuard {
    int prev = -1; bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

bool is_device_ptr(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

const swb_scoring kDefaultScoring = {3, -3, -2};     // omp_smithW.c:75-77

int check_scoring(const swb_scoring& sc, int64_t m, int64_t n)
{
    const int64_t lim = 1 << 20;
    if (std::llabs((long long)sc.match) > lim || std::llabs((long long)sc.mismatch) > lim ||
        std::llabs((long long)sc.gap) > lim)
        return SWB_ERR_RANGE;
    // a gap must cost something and a mismatch must not pay: the fill relies on H = 0 / NONE being what a cell with
    // all-zero neighbours and a non-matching character evaluates to (with gap >= 0 the scores are not bounded by
    // match * min(m,n) either)
    if (sc.gap >= 0 || sc.mismatch > 0) return SWB_ERR_RANGE;
    // packed keys are 16*H + tie in int32; H <= match * min(m,n)
    const int64_t hmax = (int64_t)std::max(sc.match, 0) * std::min(m, n);
    if (hmax >= (1LL << 26)) return SWB_ERR_RANGE;
    return SWB_OK;
}

}  // namespace
struct swb_timer { cudaEvent_t start = nullptr, stop = nullptr; int device = 0; };
namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// The workspace comes from a stream-ordered pool on every call.  The library owns one PRIVATE pool per device
// (the process's default pool and its release threshold are left alone): freed blocks stay in it up to
// kPoolKeepBytes, because handing ~100 MB of band-boundary rows back to the driver at every synchronisation
// (release threshold 0) costs milliseconds per call.  This table is the library's only process-wide state.
constexpr unsigned long long kPoolKeepBytes = 2ull << 30;
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[64] = {};

cudaMemPool_t workspace_pool(int device)
{
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
            unsigned long long keep = kPoolKeepBytes;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            g_pools[device] = pool;
        }
        cudaGetLastError();
    }
    return g_pools[device];
}

struct StripLink {                         // column-strip mode (nullptr members = not used)
    const int32_t* left_in = nullptr; const int* left_flags = nullptr;
    int32_t* right_out = nullptr; int* right_flags = nullptr; int epoch = 0;
};

}  // namespace

// Two library-owned streams per device for swb_fill_pairs_async (created on first use, kept for the life of the
// process like the workspace pool).  Streams created and destroyed inside a call did not run concurrently with each
// other (two 45000 x 45000 pairs: 11.3 ms against 7.8 ms on long-lived streams).
static bool side_streams(int device, cudaStream_t out[2])
{
    static std::mutex mu;
    static cudaStream_t table[64][2] = {};
    if (device < 0 || device >= 64) return false;
    std::lock_guard<std::mutex> lock(mu);
    for (int q = 0; q < 2; ++q)
        if (!table[device][q] && cudaStreamCreateWithFlags(&table[device][q], cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError(); table[device][q] = nullptr; return false;
        }
    out[0] = table[device][0]; out[1] = table[device][1];
    return true;
}

// The fill's host side, once per kernel geometry (see swb_fill_impl.inc)
namespace swb {
#include "swb_fill_impl.inc"
}
namespace swb_tall {
#define SWB_SINGLE_ONLY 1
#include "swb_fill_impl.inc"
#undef SWB_SINGLE_ONLY
}
namespace swb_wide {
#define SWB_SINGLE_ONLY 1
#include "swb_fill_impl.inc"
#undef SWB_SINGLE_ONLY
}

namespace {
// The one implementation behind swb_fill_async, swb_fill_batch_async and swb_score_only_async.
//   npairs equally shaped pairs: a = npairs*m bytes, b = npairs*n bytes, pair k's matrices at
//   dH/dP + k*pair_stride; d_maxPos / d_maxScore hold npairs entries.  store == false: score only.
// Single pairs (full fill and score only) run the three-rows-per-lane geometry (96-row strips: a third fewer links in
// the strip-to-strip chain that bounds large fills); batches keep two rows per lane (two or more CTAs per SM).
// The half skew of swb_tall trades a 35 % longer step (two shuffle rounds) for a 45 % shorter strip-to-strip lag; a
// matrix more than three times wider than tall has few strips and long rows, so the step dominates: full skew there
// (measured, 2 000 000 columns x 1000 rows: 52.5 ms full skew, 69.4 ms half skew; 1000 x 2 000 000: 112 ms vs 75 ms).
int fill_impl(const char* a, int64_t m, const char* b, int64_t n, int64_t npairs, const swb_scoring* scoring,
              int32_t* dH, int32_t* dP, int64_t pitch, int64_t pair_stride, int64_t* d_maxPos, int32_t* d_maxScore,
              int device, void* stream, const swb_tuning* tuning, bool store, const StripLink* link = nullptr)
{
    if (npairs == 1 && m > 3 * n)
        return swb_wide::fill_impl(a, m, b, n, npairs, scoring, dH, dP, pitch, pair_stride, d_maxPos, d_maxScore, device, stream,
                                   tuning, store, link);
    if (npairs == 1)
        return swb_tall::fill_impl(a, m, b, n, npairs, scoring, dH, dP, pitch, pair_stride, d_maxPos, d_maxScore, device, stream,
                                   tuning, store, link);
    return swb::fill_impl(a, m, b, n, npairs, scoring, dH, dP, pitch, pair_stride, d_maxPos, d_maxScore, device, stream,
                          tuning, store, link);
}
}  // namespace

extern "C" {

const char* swb_strerror(int status)
{
    switch (status) {
    case SWB_OK:        return "ok";
    case SWB_ERR_ARG:   return "invalid argument";
    case SWB_ERR_ALIGN: return "H/P device buffers must be 16-byte aligned";
    case SWB_ERR_CUDA:  return "CUDA error (see swb_last_cuda_error)";
    case SWB_ERR_RANGE: return "sizes or scores exceed the 32-bit packed score range";
    case SWB_ERR_NOMEM: return "out of device or pinned host memory";
    case SWB_ERR_IO:    return "sequence file or manifest unreadable or malformed";
    default:            return "unknown swb status";
    }
}

const char* swb_last_cuda_error(void) { return g_cuda_err; }
int swb_version(void) { return 110; }

int swb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void swb_generate(unsigned seed, int64_t m, int64_t n, char* a, char* b)
{
    // omp_smithW.c:489-519 with the padded loop bounds of :109-110 (m+1 and n+1 draws)
    auto draw = []() -> char {
        const int v = rand() % 4;
        return v == 0 ? 'A' : v == 2 ? 'C' : v == 3 ? 'G' : 'T';
    };
    srand(seed);
    for (int64_t k = 0; k <= m; ++k) { const char c = draw(); if (k < m) a[k] = c; }
    for (int64_t k = 0; k <= n; ++k) { const char c = draw(); if (k < n) b[k] = c; }
}

void* swb_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void swb_host_free(void* p) { if (p) cudaFreeHost(p); }

int swb_timer_create(swb_timer** out, int device)
{
    if (!out) return SWB_ERR_ARG;
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __LINE__);
    swb_timer* t = new swb_timer();
    t->device = device;
    cudaError_t e = cudaEventCreate(&t->start);
    if (e == cudaSuccess) e = cudaEventCreate(&t->stop);
    if (e != cudaSuccess) { swb_timer_destroy(t); return cuda_fail(e, "cudaEventCreate", __LINE__); }
    *out = t;
    return SWB_OK;
}

int swb_timer_elapsed_ms(swb_timer* t, float* ms)
{
    if (!t || !ms) return SWB_ERR_ARG;
    DeviceGuard guard(t->device);
    SWB_CUDA(cudaEventSynchronize(t->stop));
    SWB_CUDA(cudaEventElapsedTime(ms, t->start, t->stop));
    return SWB_OK;
}

void swb_timer_destroy(swb_timer* t)
{
    if (!t) return;
    DeviceGuard guard(t->device);
    if (t->start) cudaEventDestroy(t->start);
    if (t->stop) cudaEventDestroy(t->stop);
    delete t;
}

int swb_fill_async(const char* a, int64_t m, const char* b, int64_t n,
                   const swb_scoring* scoring, int32_t* dH, int32_t* dP, int64_t pitch,
                   int64_t* d_maxPos, int32_t* d_maxScore, int device, void* stream,
                   const swb_tuning* tuning)
{
    return fill_impl(a, m, b, n, 1, scoring, dH, dP, pitch, 0, d_maxPos, d_maxScore, device, stream, tuning, true);
}

int swb_fill_batch_async(const char* a, int64_t m, const char* b, int64_t n, int64_t npairs,
                         const swb_scoring* scoring, int32_t* dH, int32_t* dP, int64_t pitch, int64_t pair_stride,
                         int64_t* d_maxPos, int32_t* d_maxScore, int device, void* stream, const swb_tuning* tuning)
{
    return fill_impl(a, m, b, n, npairs, scoring, dH, dP, pitch, pair_stride, d_maxPos, d_maxScore, device, stream,
                     tuning, true);
}

int swb_fill_pairs_async(const char* a, const int64_t* a_off, const int64_t* m,
                         const char* b, const int64_t* b_off, const int64_t* n,
                         const int64_t* hp_off, int64_t npairs, const swb_scoring* scoring,
                         int32_t* dH, int32_t* dP, int64_t* d_maxPos, int32_t* d_maxScore, int device, void* stream)
{
    if (!a || !a_off || !m || !b || !b_off || !n || !hp_off || npairs <= 0 || !dH || !dP) return SWB_ERR_ARG;
    // Large pairs (one pair alone covers at least half of the GPU's 296 strip slots) run one by one on the single-pair
    // geometry, alternating between two internal streams: the first strips of pair k+1 take the SMs that the ramp-down
    // of pair k's wavefront leaves idle (measured, 45000 x 45000: 4.60 -> 3.66 ms per pair = 68 % of the HBM peak
    // sustained).  The internal streams fork from and join into the caller's stream, so the call stays asynchronous.
    auto large = [&](int64_t i) { return n[i] >= 148 * 96 && m[i] >= 4096; };
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    int64_t nlarge = 0;
    for (int64_t i = 0; i < npairs; ++i) nlarge += large(i) ? 1 : 0;
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __LINE__);
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
    auto cleanup = [&]() {
        for (int q = 0; q < 2; ++q) if (join[q]) cudaEventDestroy(join[q]);
        if (fork) cudaEventDestroy(fork);
    };
    if (nlarge >= 2 && side_streams(device, side)) {
        cudaError_t e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
        for (int q = 0; q < 2 && e == cudaSuccess; ++q) e = cudaEventCreateWithFlags(&join[q], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(fork, user);
        for (int q = 0; q < 2 && e == cudaSuccess; ++q) e = cudaStreamWaitEvent(side[q], fork, 0);
        if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "swb_fill_pairs_async streams", __LINE__); }
    } else {
        side[0] = side[1] = nullptr;
    }
    int rc = SWB_OK;
    int64_t k = 0, large_seen = 0;
    while (k < npairs && rc == SWB_OK) {
        if (hp_off[k] & 3) { rc = SWB_ERR_ALIGN; break; }
        if (large(k)) {
            cudaStream_t st = side[0] ? side[large_seen++ & 1] : user;
            rc = fill_impl(a + a_off[k], m[k], b + b_off[k], n[k], 1, scoring, dH + hp_off[k], dP + hp_off[k], m[k] + 1, 0,
                           d_maxPos ? d_maxPos + k : nullptr, d_maxScore ? d_maxScore + k : nullptr, device, st, nullptr, true);
            ++k;
            continue;
        }
        // the longest run of equally shaped (small) pairs with packed sequences and a uniform matrix stride: one launch
        int64_t run = 1, stride = 0;
        if (k + 1 < npairs) stride = hp_off[k + 1] - hp_off[k];
        while (k + run < npairs && !large(k + run) && m[k + run] == m[k] && n[k + run] == n[k] &&
               a_off[k + run] - a_off[k + run - 1] == m[k] && b_off[k + run] - b_off[k + run - 1] == n[k] &&
               hp_off[k + run] - hp_off[k + run - 1] == stride && stride >= (n[k] + 1) * (m[k] + 1))
            ++run;
        rc = fill_impl(a + a_off[k], m[k], b + b_off[k], n[k], run, scoring, dH + hp_off[k], dP + hp_off[k],
                       m[k] + 1, run > 1 ? stride : 0, d_maxPos ? d_maxPos + k : nullptr,
                       d_maxScore ? d_maxScore + k : nullptr, device, user, nullptr, true);
        k += run;
    }
    if (side[0]) {
        // join (also after an error: whatever was enqueued on the side streams must not outlive the call's ordering)
        for (int q = 0; q < 2; ++q)
            if (cudaEventRecord(join[q], side[q]) != cudaSuccess || cudaStreamWaitEvent(user, join[q], 0) != cudaSuccess) {
                if (rc == SWB_OK) rc = cuda_fail(cudaGetLastError(), "swb_fill_pairs_async join", __LINE__);
            }
    }
    cleanup();
    return rc;
}

int swb_shard_pairs(int64_t npairs, int nshards, int shard, int64_t* first, int64_t* count)
{
    if (npairs < 0 || nshards < 1 || shard < 0 || shard >= nshards || !first || !count) return SWB_ERR_ARG;
    const int64_t base = npairs / nshards, rem = npairs % nshards;
    *first = base * shard + std::min<int64_t>(shard, rem);
    *count = base + (shard < rem ? 1 : 0);
    return SWB_OK;
}

int swb_fill_strip_async(const char* a_local, int64_t m_local, const char* b, int64_t n,
                         const swb_scoring* scoring, int32_t* dH, int32_t* dP, int64_t pitch,
                         const int32_t* left_in, const int32_t* left_flags, int32_t* right_out, int32_t* right_flags,
                         int32_t epoch, int64_t* d_maxPos, int32_t* d_maxScore, int device, void* stream,
                         const swb_tuning* tuning)
{
    if ((left_in == nullptr) != (left_flags == nullptr) || (right_out == nullptr) != (right_flags == nullptr) || epoch == 0)
        return SWB_ERR_ARG;
    StripLink link;
    link.left_in = left_in; link.left_flags = left_flags; link.right_out = right_out; link.right_flags = right_flags;
    link.epoch = epoch;
    return fill_impl(a_local, m_local, b, n, 1, scoring, dH, dP, pitch, 0, d_maxPos, d_maxScore, device, stream, tuning, true,
                     &link);
}

int64_t swb_strip_flag_count(int64_t n) { return n <= 0 ? 0 : (n + swb::kWRows - 1) / swb::kWRows; }

void* swb_ipc_alloc(size_t bytes, int device)
{
    DeviceGuard guard(device);
    void* p = nullptr;
    if (!guard.ok || cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemset(p, 0, bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(p); return nullptr; }
    return p;
}
void swb_ipc_free(void* p, int device) { DeviceGuard guard(device); if (p) cudaFree(p); }
int swb_ipc_get_handle(void* p, unsigned char* out64)
{
    if (!p || !out64) return SWB_ERR_ARG;
    cudaIpcMemHandle_t h;
    SWB_CUDA(cudaIpcGetMemHandle(&h, p));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(out64, &h, 64);
    return SWB_OK;
}
int swb_ipc_open(const unsigned char* handle64, int device, void** out)
{
    if (!handle64 || !out) return SWB_ERR_ARG;
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __LINE__);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    SWB_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return SWB_OK;
}
int swb_ipc_close(void* p, int device)
{
    DeviceGuard guard(device);
    if (p) SWB_CUDA(cudaIpcCloseMemHandle(p));
    return SWB_OK;
}
int swb_enable_peer(int device, int peer)
{
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __LINE__);
    int can = 0;
    SWB_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
    if (!can) return SWB_ERR_ARG;
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __LINE__);
    cudaGetLastError();
    return SWB_OK;
}

int swb_score_only_async(const char* a, int64_t m, const char* b, int64_t n, int64_t npairs,
                         const swb_scoring* scoring, int64_t* d_maxPos, int32_t* d_maxScore,
                         int device, void* stream, const swb_tuning* tuning)
{
    return fill_impl(a, m, b, n, npairs, scoring, nullptr, nullptr, m + 1, 0, d_maxPos, d_maxScore, device, stream,
                     tuning, false);
}

int swb_fill(const char* a, int64_t m, const char* b, int64_t n,
             const swb_scoring* scoring, int32_t* dH, int32_t* dP, int64_t pitch,
             int64_t* maxPos, int device, void* stream)
{
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __LINE__);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long* d_pos = nullptr;
    SWB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_pos), sizeof(long long), st));
    int rc = swb_fill_async(a, m, b, n, scoring, dH, dP, pitch, reinterpret_cast<int64_t*>(d_pos), nullptr,
                            device, stream, nullptr);
    long long pos = 0;
    if (rc == SWB_OK) {
        cudaError_t e = cudaMemcpyAsync(&pos, d_pos, sizeof pos, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "maxPos readback", __LINE__);
    }
    cudaFreeAsync(d_pos, st);
    if (rc == SWB_OK && maxPos) *maxPos = pos;
    return rc;
}

}  // extern "C"
namespace {
// the one backtrack launch: start cell from d_maxPos (device) or maxPos; optional path length, end cell and moves
int launch_backtrack(int32_t* dP, int64_t pitch, int64_t maxPos, const int64_t* d_maxPos, int64_t* d_pathLen,
                     int64_t* d_endPos, unsigned char* d_moves, int device, void* stream)
{
    if (!dP || pitch <= 1 || (!d_maxPos && maxPos < 0)) return SWB_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(dP) & 15) return SWB_ERR_ALIGN;      // band rows are fetched with 16-byte bulk copies
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __LINE__);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (SWB_BT2 || d_moves) {
        SWB_CUDA(cudaFuncSetAttribute(swb::bt2::backtrack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, swb::bt2::kSmemBytes));
        swb::bt2::backtrack_kernel<<<1, swb::bt2::kThreads, swb::bt2::kSmemBytes, st>>>(
            dP, pitch, maxPos, reinterpret_cast<const long long*>(d_maxPos), reinterpret_cast<long long*>(d_pathLen),
            reinterpret_cast<long long*>(d_endPos), d_moves);
        SWB_CUDA(cudaGetLastError());
        return SWB_OK;
    }
    const int bt_smem = (2 * (swb::kBtPad + swb::kBtBandInts) + 2 * swb::kBtList) * (int)sizeof(int);
    SWB_CUDA(cudaFuncSetAttribute(swb::backtrack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bt_smem));
    swb::backtrack_kernel<<<1, swb::kBtThreads, bt_smem, st>>>(dP, pitch, maxPos, reinterpret_cast<const long long*>(d_maxPos),
                                                               reinterpret_cast<long long*>(d_pathLen),
                                                               reinterpret_cast<long long*>(d_endPos));
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}
}  // namespace
extern "C" {

int swb_backtrack_from_async(int32_t* dP, int64_t pitch, int64_t startPos, int64_t* d_pathLen, int64_t* d_endPos,
                             int device, void* stream)
{
    return launch_backtrack(dP, pitch, startPos, nullptr, d_pathLen, d_endPos, nullptr, device, stream);
}

int swb_backtrack_async(int32_t* dP, int64_t pitch, int64_t maxPos, const int64_t* d_maxPos,
                        int64_t* d_pathLen, int device, void* stream)
{
    return launch_backtrack(dP, pitch, maxPos, d_maxPos, d_pathLen, nullptr, nullptr, device, stream);
}

int swb_traceback_async(int32_t* dP, int64_t pitch, int64_t startPos, const int64_t* d_startPos, int64_t* d_pathLen,
                        int64_t* d_endPos, unsigned char* d_moves, int device, void* stream)
{
    return launch_backtrack(dP, pitch, startPos, d_startPos, d_pathLen, d_endPos, d_moves, device, stream);
}

// moves in walk order (from the start cell backwards) -> CIGAR in sequence order.  a (columns) is the reference,
// b (rows) the query: DIAGONAL consumes both (M), UP consumes a row = query only (I), LEFT a column = reference only (D).
int64_t swb_cigar_from_moves(const unsigned char* moves, int64_t n, char* cigar, size_t cap)
{
    if (!moves || n < 0 || (!cigar && cap)) return SWB_ERR_ARG;
    size_t used = 0;
    int64_t k = n - 1;
    while (k >= 0) {
        const unsigned char mv = moves[k];
        int64_t run = 0;
        while (k >= 0 && moves[k] == mv) { ++run; --k; }
        const char op = mv == 3 ? 'M' : mv == 1 ? 'I' : mv == 2 ? 'D' : '?';
        char buf[32];
        const int len = std::snprintf(buf, sizeof buf, "%lld%c", (long long)run, op);
        if (cigar && used + (size_t)len < cap) std::memcpy(cigar + used, buf, (size_t)len);
        used += (size_t)len;
    }
    if (cigar && cap) cigar[used < cap ? used : cap - 1] = 0;
    return (int64_t)used;                                                 // length needed (without the NUL)
}

// the two gapped strings of the alignment ('-' = gap), sequence order; out_a / out_b hold n + 1 bytes.
// startPos = the cell the walk started from (maxPos), pitch = ints per row of the matrix it indexes.
int swb_alignment_from_moves(const unsigned char* moves, int64_t n, const char* a, const char* b, int64_t startPos,
                             int64_t pitch, char* out_a, char* out_b)
{
    if (!moves || n < 0 || !a || !b || !out_a || !out_b || pitch <= 1) return SWB_ERR_ARG;
    int64_t i = startPos / pitch, j = startPos % pitch;
    for (int64_t k = 0; k < n; ++k) {
        const int64_t o = n - 1 - k;
        if (i < 1 && moves[k] != 2) return SWB_ERR_ARG;
        if (j < 1 && moves[k] != 1) return SWB_ERR_ARG;
        switch (moves[k]) {                                               // cell (i, j) pairs b[i-1] with a[j-1] (omp_smithW.c:395)
        case 3: out_a[o] = a[j - 1]; out_b[o] = b[i - 1]; --i; --j; break;
        case 1: out_a[o] = '-';      out_b[o] = b[i - 1]; --i; break;
        case 2: out_a[o] = a[j - 1]; out_b[o] = '-';      --j; break;
        default: return SWB_ERR_ARG;
        }
    }
    out_a[n] = 0; out_b[n] = 0;
    return SWB_OK;
}

int swb_backtrack(int32_t* dP, int64_t pitch, int64_t maxPos, int64_t* path_len, int device, void* stream)
{
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __LINE__);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long* d_len = nullptr;
    SWB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_len), sizeof(long long), st));
    int rc = swb_backtrack_async(dP, pitch, maxPos, nullptr, reinterpret_cast<int64_t*>(d_len), device, stream);
    long long len = 0;
    if (rc == SWB_OK) {
        cudaError_t e = cudaMemcpyAsync(&len, d_len, sizeof len, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "path length readback", __LINE__);
    }
    cudaFreeAsync(d_len, st);
    if (rc == SWB_OK && path_len) *path_len = len;
    return rc;
}

// ------------------------------------------------------------------ host-buffer API
struct swb_ctx {
    int device = 0;
    int64_t m = 0, n = 0;
    int32_t* dH = nullptr; int32_t* dP = nullptr;
    long long* d_scalars = nullptr;        // [0] maxPos, [1] pathLen, [2] packed-transfer overflow flag
    cudaStream_t st = nullptr;
    // packed copy-back (swb_pack.cu): scratch for swb_d2h_packed on the device and in pinned host memory, allocated on first use
    unsigned char* d_packed = nullptr; unsigned char* h_packed = nullptr;
    bool packed_tried = false;
};


int swb_ctx_create(swb_ctx** out, int64_t m, int64_t n, int device)
{
    if (!out || m <= 0 || n <= 0) return SWB_ERR_ARG;
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __LINE__);
    swb_ctx* c = new swb_ctx();
    c->device = device; c->m = m; c->n = n;
    const size_t bytes = (size_t)(m + 1) * (size_t)(n + 1) * sizeof(int32_t);
    cudaError_t e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->dH), bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->dP), bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->d_scalars), 3 * sizeof(long long));
    if (e != cudaSuccess) { int rc = cuda_fail(e, "swb_ctx_create allocation", __LINE__); swb_ctx_destroy(c); return rc; }
    *out = c;
    return SWB_OK;
}

void swb_ctx_destroy(swb_ctx* c)
{
    if (!c) return;
    DeviceGuard guard(c->device);
    if (c->dH) cudaFree(c->dH);
    if (c->dP) cudaFree(c->dP);
    if (c->d_scalars) cudaFree(c->d_scalars);
    if (c->d_packed) cudaFree(c->d_packed);
    if (c->h_packed) cudaFreeHost(c->h_packed);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

int32_t* swb_ctx_dH(swb_ctx* c) { return c ? c->dH : nullptr; }
int32_t* swb_ctx_dP(swb_ctx* c) { return c ? c->dP : nullptr; }

// Packed copy-back is used for matrices of at least this many bytes each (below it the plain copies take well under a
// millisecond); SWB_PACKED_D2H=0 turns it off, =1 forces it for every size.
static bool packed_wanted(size_t bytes)
{
    static const int mode = [] { const char* s = std::getenv("SWB_PACKED_D2H"); return s ? std::atoi(s) : -1; }();
    if (mode == 0) return false;
    if (mode > 0) return true;
    return bytes >= ((size_t)32 << 20);
}

int swb_ctx_align(swb_ctx* c, const char* a, const char* b, const swb_scoring* scoring,
                  int32_t* H, int32_t* P, int64_t* maxPos, int64_t* path_len, int do_backtrack)
{
    if (!c || !a || !b) return SWB_ERR_ARG;
    DeviceGuard guard(c->device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __LINE__);
    const size_t bytes = (size_t)(c->m + 1) * (size_t)(c->n + 1) * sizeof(int32_t);
    bool packed = (H || P) && packed_wanted(bytes);
    if (packed && !c->packed_tried) {
        // (a context whose packed buffers cannot be had keeps working with the plain copies)
        c->packed_tried = true;
        const size_t pbytes = swb_d2h_packed_scratch_bytes(c->n + 1, c->m + 1);
        if (cudaMalloc(reinterpret_cast<void**>(&c->d_packed), pbytes) != cudaSuccess) { c->d_packed = nullptr; cudaGetLastError(); }
        else if (cudaHostAlloc(reinterpret_cast<void**>(&c->h_packed), pbytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError(); cudaFree(c->d_packed); c->d_packed = nullptr; c->h_packed = nullptr;
        }
    }
    packed = packed && c->d_packed && c->h_packed;
    int rc = swb_fill_async(a, c->m, b, c->n, scoring, c->dH, c->dP, c->m + 1,
                            reinterpret_cast<int64_t*>(c->d_scalars), nullptr, c->device, c->st, nullptr);
    if (rc != SWB_OK) return rc;
    // plain copies: H does not change any more, start its copy-back before the backtrack
    if (H && !packed) SWB_CUDA(cudaMemcpyAsync(H, c->dH, bytes, cudaMemcpyDeviceToHost, c->st));
    if (do_backtrack) {
        rc = swb_backtrack_async(c->dP, c->m + 1, 0, reinterpret_cast<int64_t*>(c->d_scalars),
                                 reinterpret_cast<int64_t*>(c->d_scalars + 1), c->device, c->st);
        if (rc != SWB_OK) return rc;
    } else {
        SWB_CUDA(cudaMemsetAsync(c->d_scalars + 1, 0, sizeof(long long), c->st));
    }
    long long sc3[3] = {0, 0, 0};
    SWB_CUDA(cudaMemcpyAsync(sc3, c->d_scalars, 2 * sizeof(long long), cudaMemcpyDeviceToHost, c->st));
    if (packed) {
        // one byte per cell over PCIe, expanded into the caller's int32 matrices on the host (swb_pack.cu)
        rc = swb_d2h_packed(c->dH, c->dP, c->m + 1, c->n + 1, c->m + 1, H, P, c->m + 1, c->d_packed, c->h_packed, 0, c->device, c->st);
        if (rc != SWB_OK) return rc == SWB_ERR_CUDA ? cuda_fail(cudaGetLastError(), "swb_d2h_packed", __LINE__) : rc;
    } else {
        if (P) SWB_CUDA(cudaMemcpyAsync(P, c->dP, bytes, cudaMemcpyDeviceToHost, c->st));
        SWB_CUDA(cudaStreamSynchronize(c->st));
    }
    if (maxPos) *maxPos = sc3[0];
    if (path_len) *path_len = sc3[1];
    return SWB_OK;
}

int swb_align_host(const char* a, int64_t m, const char* b, int64_t n,
                   const swb_scoring* scoring, int32_t* H, int32_t* P,
                   int64_t* maxPos, int64_t* path_len, int do_backtrack, int device)
{
    swb_ctx* c = nullptr;
    int rc = swb_ctx_create(&c, m, n, device);
    if (rc != SWB_OK) return rc;
    rc = swb_ctx_align(c, a, b, scoring, H, P, maxPos, path_len, do_backtrack);
    swb_ctx_destroy(c);
    return rc;
}

int swb_score_only(const char* a, int64_t m, const char* b, int64_t n,
                   const swb_scoring* scoring, int32_t* maxScore, int64_t* maxPos,
                   int device, void* stream)
{
    if (m <= 0 || n <= 0) return SWB_ERR_ARG;
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __LINE__);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long* d_pos = nullptr;
    SWB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_pos), 16, st));
    int32_t* d_sc = reinterpret_cast<int32_t*>(d_pos + 1);
    int rc = swb_score_only_async(a, m, b, n, 1, scoring, reinterpret_cast<int64_t*>(d_pos), d_sc, device, stream, nullptr);
    long long host[2] = {0, 0};
    if (rc == SWB_OK) {
        cudaError_t e = cudaMemcpyAsync(host, d_pos, 16, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "score readback", __LINE__);
    }
    cudaFreeAsync(d_pos, st);
    if (rc == SWB_OK) {
        if (maxPos) *maxPos = host[0];
        if (maxScore) *maxScore = (int32_t)(host[1] & 0xffffffff);
    }
    return rc;
}

}  // extern "C"
