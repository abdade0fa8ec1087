// swb_multi.cu -- C++ host driver for ONE pair split into column strips over the GPUs of one box
// (BASELINE config "100000x100000 single pair, column-strip wavefront pipelined across 2/4/8 B200 with
// NVLink P2P boundary exchange"; SURVEY 8(b) swb_fill_multi, 8(e)).  One process, one stream per GPU.
//
// The reference has no multi-GPU form; this extends its nDiag wavefront (omp_smithW.c:203-216): GPU g owns a
// contiguous block of columns of every row as its own row-major slab and its fill kernel pushes the H values of
// its last column into GPU g+1's memory with peer stores while it runs (swb_fill_strip_async).  Nothing but those
// stores crosses the GPUs during the fill; afterwards the host gathers one (score, position) pair per GPU for
// maxPos (reference tie-break, omp_smithW.c:384-387) and walks the backtrack right to left over the strips
// (omp_smithW.c:405-420).
#include "../../include/swb200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

struct swb_multi {
    struct Strip {
        int device = 0;
        int64_t col0 = 0, m = 0, pitch = 0;              // global columns col0+1 .. col0+m; local column 0 = column col0
        char* a_d = nullptr; char* b_d = nullptr;
        int32_t* dH = nullptr; int32_t* dP = nullptr;
        int32_t* left_in[2] = {nullptr, nullptr};        // double-buffered by call parity (see ADVICE: no reuse before
        int32_t* left_flags[2] = {nullptr, nullptr};     // the previous call has been synchronised)
        int64_t* d_scal = nullptr;                       // [0] maxPos (local), [1] path length, [2] end of a walk
        int32_t* d_score = nullptr;
        cudaStream_t st = nullptr;
        bool own_stream = false;                         // strips that share a device share a stream: they run left to
                                                         // right in stream order (two kernels of one device could starve
                                                         // each other's waiting CTAs otherwise)
    };
    int64_t m = 0, n = 0;
    int epoch = 0;
    std::vector<Strip> strips;
};

namespace {

struct Guard {
    int prev = -1;
    explicit Guard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; cudaSetDevice(dev); }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define SWBM_CUDA(call) do { if ((call) != cudaSuccess) { cudaGetLastError(); return SWB_ERR_CUDA; } } while (0)

}  // namespace

extern "C" {

int swb_multi_create(swb_multi** out, int64_t m, int64_t n, const int* devices, int ndev)
{
    if (!out || !devices || ndev < 1 || m < ndev || n <= 0) return SWB_ERR_ARG;
    swb_multi* h = new swb_multi();
    h->m = m; h->n = n;
    h->strips.resize((size_t)ndev);
    const int64_t base = m / ndev, rem = m % ndev;
    int64_t c = 0;
    int rc = SWB_OK;
    for (int g = 0; g < ndev && rc == SWB_OK; ++g) {
        swb_multi::Strip& s = h->strips[(size_t)g];
        s.device = devices[g];
        s.col0 = c; s.m = base + (g < rem ? 1 : 0); s.pitch = s.m + 1;
        c += s.m;
        Guard guard(s.device);
        const size_t cells = (size_t)(n + 1) * (size_t)s.pitch;
        const size_t nflags = (size_t)std::max<int64_t>(swb_strip_flag_count(n), 1);
        cudaError_t e = cudaSuccess;
        for (int q = 0; q < g; ++q)
            if (h->strips[(size_t)q].device == s.device) { s.st = h->strips[(size_t)q].st; break; }
        if (!s.st) { e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking); s.own_stream = (e == cudaSuccess); }
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.a_d, (size_t)s.m);
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.b_d, (size_t)n);
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.dH, cells * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.dP, cells * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.d_scal, 3 * sizeof(int64_t));
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.d_score, sizeof(int32_t));
        for (int k = 0; k < 2 && g > 0 && e == cudaSuccess; ++k) {
            e = cudaMalloc((void**)&s.left_in[k], (size_t)(n + 1) * sizeof(int32_t));
            if (e == cudaSuccess) e = cudaMalloc((void**)&s.left_flags[k], nflags * sizeof(int32_t));
            if (e == cudaSuccess) e = cudaMemset(s.left_flags[k], 0, nflags * sizeof(int32_t));
        }
        if (e != cudaSuccess) { cudaGetLastError(); rc = (e == cudaErrorMemoryAllocation) ? SWB_ERR_NOMEM : SWB_ERR_CUDA; }
    }
    // GPU g stores into GPU g+1's boundary buffers
    for (int g = 0; g + 1 < ndev && rc == SWB_OK; ++g)
        if (devices[g] != devices[g + 1]) rc = swb_enable_peer(devices[g], devices[g + 1]);
    if (rc != SWB_OK) { swb_multi_destroy(h); return rc; }
    *out = h;
    return SWB_OK;
}

void swb_multi_destroy(swb_multi* h)
{
    if (!h) return;
    for (auto& s : h->strips) { Guard guard(s.device); if (s.st) cudaStreamSynchronize(s.st); }
    for (auto& s : h->strips) {
        Guard guard(s.device);
        if (s.st && s.own_stream) cudaStreamDestroy(s.st);
        cudaFree(s.a_d); cudaFree(s.b_d); cudaFree(s.dH); cudaFree(s.dP); cudaFree(s.d_scal); cudaFree(s.d_score);
        for (int k = 0; k < 2; ++k) { cudaFree(s.left_in[k]); cudaFree(s.left_flags[k]); }
    }
    cudaGetLastError();
    delete h;
}

int swb_multi_strips(const swb_multi* h) { return h ? (int)h->strips.size() : 0; }

int swb_multi_strip(const swb_multi* h, int g, int* device, int64_t* col0, int64_t* m_local, int64_t* pitch,
                    int32_t** dH, int32_t** dP)
{
    if (!h || g < 0 || g >= (int)h->strips.size()) return SWB_ERR_ARG;
    const swb_multi::Strip& s = h->strips[(size_t)g];
    if (device) *device = s.device;
    if (col0) *col0 = s.col0;
    if (m_local) *m_local = s.m;
    if (pitch) *pitch = s.pitch;
    if (dH) *dH = s.dH;
    if (dP) *dP = s.dP;
    return SWB_OK;
}

// fill (all strips at once) + maxPos.  a, b: HOST sequences of the whole pair.
int swb_multi_fill(swb_multi* h, const char* a, const char* b, const swb_scoring* scoring,
                   int64_t* maxPos, int32_t* maxScore)
{
    if (!h || !a || !b) return SWB_ERR_ARG;
    const int G = (int)h->strips.size();
    const int64_t m = h->m, n = h->n;
    // every stream was synchronised at the end of the previous call, so the boundary buffers of parity
    // (epoch & 1) are free again: a left GPU can never run ahead of a right one by more than this call
    h->epoch += 1;
    if (h->epoch <= 0) h->epoch = 1;
    const int k = h->epoch & 1;
    for (int g = 0; g < G; ++g) {                         // left to right: a strip's producer is always enqueued first
        swb_multi::Strip& s = h->strips[(size_t)g];
        Guard guard(s.device);
        SWBM_CUDA(cudaMemcpyAsync(s.a_d, a + s.col0, (size_t)s.m, cudaMemcpyHostToDevice, s.st));
        SWBM_CUDA(cudaMemcpyAsync(s.b_d, b, (size_t)n, cudaMemcpyHostToDevice, s.st));
        const swb_multi::Strip* r = (g + 1 < G) ? &h->strips[(size_t)g + 1] : nullptr;
        const int rc = swb_fill_strip_async(s.a_d, s.m, s.b_d, n, scoring, s.dH, s.dP, s.pitch,
                                            g > 0 ? s.left_in[k] : nullptr, g > 0 ? s.left_flags[k] : nullptr,
                                            r ? r->left_in[k] : nullptr, r ? r->left_flags[k] : nullptr, h->epoch,
                                            s.d_scal, s.d_score, s.device, s.st, nullptr);
        if (rc != SWB_OK) return rc;
    }
    // maxPos: first cell in the reference's scan order (smallest i+j, then largest i) among the strips' maxima
    int64_t best_pos = 0, best_i = 0, best_j = 0; int32_t best_score = 0;
    for (int g = 0; g < G; ++g) {
        swb_multi::Strip& s = h->strips[(size_t)g];
        Guard guard(s.device);
        int64_t pos = 0; int32_t score = 0;
        SWBM_CUDA(cudaMemcpyAsync(&pos, s.d_scal, sizeof pos, cudaMemcpyDeviceToHost, s.st));
        SWBM_CUDA(cudaMemcpyAsync(&score, s.d_score, sizeof score, cudaMemcpyDeviceToHost, s.st));
        SWBM_CUDA(cudaStreamSynchronize(s.st));
        if (score <= 0) continue;
        const int64_t i = pos / s.pitch, j = s.col0 + pos % s.pitch;
        const bool better = score > best_score ||
                            (score == best_score && (i + j < best_i + best_j || (i + j == best_i + best_j && i > best_i)));
        if (better) { best_score = score; best_i = i; best_j = j; best_pos = i * (m + 1) + j; }
    }
    if (maxPos) *maxPos = best_pos;
    if (maxScore) *maxScore = best_score;
    return SWB_OK;
}

// backtrack (omp_smithW.c:405-420) from maxPos (index into the (n+1) x (m+1) matrix of the whole pair)
int swb_multi_backtrack(swb_multi* h, int64_t maxPos, int64_t* path_len)
{
    if (!h || maxPos < 0) return SWB_ERR_ARG;
    const int G = (int)h->strips.size();
    const int64_t best_pos = maxPos, best_i = maxPos / (h->m + 1), best_j = maxPos % (h->m + 1);
    int64_t total = 0;
    if (best_pos > 0 && best_j >= 1) {
        // right to left: a strip's walk ends on NONE or on the hand-off marker of its local column 0, which is the
        // last column of the strip on its left
        int g = 0;
        while (g + 1 < G && best_j > h->strips[(size_t)g].col0 + h->strips[(size_t)g].m) ++g;
        int64_t i = best_i, jl = best_j - h->strips[(size_t)g].col0;
        while (true) {
            swb_multi::Strip& s = h->strips[(size_t)g];
            Guard guard(s.device);
            const int rc = swb_backtrack_from_async(s.dP, s.pitch, i * s.pitch + jl, s.d_scal + 1, s.d_scal + 2, s.device, s.st);
            if (rc != SWB_OK) return rc;
            int64_t res[2] = {0, 0};
            SWBM_CUDA(cudaMemcpyAsync(res, s.d_scal + 1, sizeof res, cudaMemcpyDeviceToHost, s.st));
            SWBM_CUDA(cudaStreamSynchronize(s.st));
            total += res[0];
            const int64_t ei = res[1] / s.pitch, ej = res[1] % s.pitch;
            if (g > 0 && ej == 0 && ei >= 1) { --g; i = ei; jl = h->strips[(size_t)g].m; }
            else break;
        }
    }
    if (path_len) *path_len = total;
    return SWB_OK;
}

int swb_multi_align(swb_multi* h, const char* a, const char* b, const swb_scoring* scoring,
                    int64_t* maxPos, int32_t* maxScore, int64_t* path_len, int do_backtrack)
{
    int64_t pos = 0;
    int rc = swb_multi_fill(h, a, b, scoring, &pos, maxScore);
    if (rc != SWB_OK) return rc;
    if (maxPos) *maxPos = pos;
    if (path_len) *path_len = 0;
    return do_backtrack ? swb_multi_backtrack(h, pos, path_len) : SWB_OK;
}

// whole-pair matrices on the host, row-major with pitch m+1 (the reference's layout, omp_smithW.c:113-118,336);
// either pointer may be NULL
int swb_multi_gather_host(swb_multi* h, int32_t* H, int32_t* P)
{
    if (!h) return SWB_ERR_ARG;
    const int64_t m = h->m, n = h->n;
    for (auto& s : h->strips) {
        Guard guard(s.device);
        // local columns 1..m_local -> global columns col0+1..; global column 0 comes from the first strip's column 0
        const int64_t first = (s.col0 == 0) ? 0 : 1;
        const size_t width = (size_t)(s.m + 1 - first) * sizeof(int32_t);
        if (H) SWBM_CUDA(cudaMemcpy2DAsync(H + s.col0 + first, (size_t)(m + 1) * 4, s.dH + first, (size_t)s.pitch * 4, width,
                                           (size_t)(n + 1), cudaMemcpyDeviceToHost, s.st));
        if (P) SWBM_CUDA(cudaMemcpy2DAsync(P + s.col0 + first, (size_t)(m + 1) * 4, s.dP + first, (size_t)s.pitch * 4, width,
                                           (size_t)(n + 1), cudaMemcpyDeviceToHost, s.st));
    }
    for (auto& s : h->strips) { Guard guard(s.device); SWBM_CUDA(cudaStreamSynchronize(s.st)); }
    return SWB_OK;
}

int swb_fill_multi(const char* a, int64_t m, const char* b, int64_t n, const swb_scoring* scoring,
                   const int* devices, int ndev, int32_t* H, int32_t* P,
                   int64_t* maxPos, int64_t* path_len, int do_backtrack)
{
    swb_multi* h = nullptr;
    int rc = swb_multi_create(&h, m, n, devices, ndev);
    if (rc != SWB_OK) return rc;
    rc = swb_multi_align(h, a, b, scoring, maxPos, nullptr, path_len, do_backtrack);
    if (rc == SWB_OK && (H || P)) rc = swb_multi_gather_host(h, H, P);
    swb_multi_destroy(h);
    return rc;
}

}  // extern "C"
