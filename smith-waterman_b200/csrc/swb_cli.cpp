// swb -- drop-in for the reference program's command line (omp_smithW.c:8,87-253):
//   ./swb <number_of_col> <number_of_rows>      random DNA of that shape
//   ./swb                                       built-in 8x9 example + self-check
// stdout keeps the reference's lines, in order, so its run scripts (which grep
// "Elapsed time for scoring matrix computation", readme.liao:12) work unchanged;
// extra lines (GCUPS, maxPos) come after them.  Additive knobs via environment:
//   SWB_SEED=<n>   pin srand() (the reference uses time(NULL), omp_smithW.c:491)
//   SWB_DEVICE=<d> CUDA device (default 0)
//   SWB_DEVICES=0,1,2,3   ONE pair in column strips over these GPUs (swb_fill_multi: NVLink boundary stores)
//   SWB_DEBUG=1    print the H and P matrices like -DDEBUG (omp_smithW.c:236-242,426-483)
//   SWB_V1_LINES=1 also print the lines the v1 variant adds, at v1's positions: "Total memory footprint is:..."
//                  (omp_smithW-v1-refinedOrig.cpp:138-142) and, with SWB_SKIP_BACKTRACK=1, "Skipping backtrack ..."
//                  (:190-192)
//   SWB_SKIP_BACKTRACK=1  the -DSKIP_BACKTRACK=1 build of v1 (makefile:9): no maxPos tracking, no backtrack; the
//                  score-only kernel runs (no H/P stores) -- the configuration all of the reference's logs timed
// Real sequences instead of generate() (additive forms of the command line):
//   ./swb --fasta <a.fa|a.2bit>[:record] <b.fa|b.2bit>[:record]     one pair from files (a = columns, b = rows)
//   ./swb --pairs <manifest>                                         a batch: one "<fileA>[:rec] <fileB>[:rec]" per line
// All compute goes through the C ABI of include/swb200.h; there is no CPU path.
#include "swb200.h"
#include "parameters.h"

#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <algorithm>
#include <string>
#include <vector>

#define RESET   "\033[0m"
#define BOLDRED "\033[1m\033[31m"

static double now_s()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

static void die(const char* what, int rc)
{
    std::fprintf(stderr, "swb: %s: %s %s\n", what, swb_strerror(rc), swb_last_cuda_error());
    std::exit(1);
}

// omp_smithW.c:426-441
static void print_matrix(const std::vector<int32_t>& M, long long cols1, long long rows1,
                         const std::string& a, const std::string& b)
{
    std::printf("-\t-\t");
    for (long long j = 0; j < cols1 - 1; j++) std::printf("%c\t", a[j]);
    std::printf("\n-\t");
    for (long long i = 0; i < rows1; i++) {
        for (long long j = 0; j < cols1; j++) {
            if (j == 0 && i > 0) std::printf("%c\t", b[i - 1]);
            std::printf("%d\t", M[cols1 * i + j]);
        }
        std::printf("\n");
    }
}

// omp_smithW.c:447-483
static void print_pred(const std::vector<int32_t>& M, long long cols1, long long rows1,
                       const std::string& a, const std::string& b)
{
    std::printf("    ");
    for (long long j = 0; j < cols1 - 1; j++) std::printf("%c ", a[j]);
    std::printf("\n  ");
    for (long long i = 0; i < rows1; i++) {
        for (long long j = 0; j < cols1; j++) {
            if (j == 0 && i > 0) std::printf("%c ", b[i - 1]);
            const int v = M[cols1 * i + j];
            const int p = v < 0 ? -v : v;
            if (v < 0) std::printf(BOLDRED);
            std::printf(p == 1 ? "↑ " : p == 2 ? "← " : p == 3 ? "↖ " : "- ");
            if (v < 0) std::printf(RESET);
        }
        std::printf("\n");
    }
}

static bool read_spec(const char* spec, std::string& out)
{
    std::string file(spec); long rec = 0;
    const size_t colon = file.find_last_of(':');
    if (colon != std::string::npos && colon + 1 < file.size() && file.find_first_not_of("0123456789", colon + 1) == std::string::npos) {
        rec = std::atol(file.c_str() + colon + 1);
        file.resize(colon);
    }
    char* seq = nullptr; int64_t len = 0;
    const int rc = swb_seq_read(file.c_str(), rec, &seq, &len, nullptr, 0);
    if (rc != SWB_OK) { std::fprintf(stderr, "swb: %s: %s\n", spec, swb_strerror(rc)); return false; }
    out.assign(seq, (size_t)len);
    swb_seq_free(seq);
    return len > 0;
}

// v1's footprint line (omp_smithW-v1-refinedOrig.cpp:138-142; its m, n are the padded sizes)
static void print_footprint(long long m, long long n)
{
    const unsigned long long M = (unsigned long long)m + 1, N = (unsigned long long)n + 1;
    const unsigned long long sz = (M + N + 2 * M * N) * sizeof(int) / 1024 / 1024;
    if (sz >= 1024) std::printf("Total memory footprint is:%llu GB\n", sz / 1024);
    else            std::printf("Total memory footprint is:%llu MB\n", sz);
}

// --pairs <manifest>: a variable-length batch through swb_fill_pairs_async + one backtrack per pair
static int run_manifest(const char* path, int device)
{
    swb_manifest* mf = nullptr;
    int rc = swb_manifest_load(path, &mf);
    if (rc != SWB_OK) { std::fprintf(stderr, "swb: %s: %s\n", path, swb_strerror(rc)); return 1; }
    const int64_t np = swb_manifest_pairs(mf);
    std::printf("Batch of %lld pairs from %s\n", (long long)np, path);
    std::vector<int64_t> a_off(np), b_off(np), mm(np), nn(np), hp_off(np);
    std::string A, B;
    int64_t cells = 0;
    for (int64_t k = 0; k < np; ++k) {
        const char *pa, *pb;
        swb_manifest_pair(mf, k, &pa, &mm[k], &pb, &nn[k]);
        a_off[k] = (int64_t)A.size(); b_off[k] = (int64_t)B.size();
        A.append(pa, (size_t)mm[k]); B.append(pb, (size_t)nn[k]);
        hp_off[k] = cells;
        cells += ((mm[k] + 1) * (nn[k] + 1) + 3) / 4 * 4;
    }
    swb_manifest_free(mf);
    cudaSetDevice(device);
    int32_t *dH = nullptr, *dP = nullptr; int64_t* d_pos = nullptr; int32_t* d_sc = nullptr;
    if (cudaMalloc((void**)&dH, (size_t)cells * 4) != cudaSuccess || cudaMalloc((void**)&dP, (size_t)cells * 4) != cudaSuccess ||
        cudaMalloc((void**)&d_pos, (size_t)np * 8) != cudaSuccess || cudaMalloc((void**)&d_sc, (size_t)np * 4) != cudaSuccess) {
        std::fprintf(stderr, "swb: cannot allocate %.2f GB of device memory for H and P\n", cells * 8 / 1e9);
        return 1;
    }
    double t0 = now_s();
    rc = swb_fill_pairs_async(A.data(), a_off.data(), mm.data(), B.data(), b_off.data(), nn.data(), hp_off.data(), np, nullptr,
                              dH, dP, d_pos, d_sc, device, nullptr);
    if (rc) die("fill", rc);
    cudaDeviceSynchronize();
    double t1 = now_s();
    std::printf("Elapsed time for scoring matrix computation: %f\n", t1 - t0);
    std::vector<int64_t> pos(np); std::vector<int32_t> sc(np);
    cudaMemcpy(pos.data(), d_pos, (size_t)np * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(sc.data(), d_sc, (size_t)np * 4, cudaMemcpyDeviceToHost);
    t0 = now_s();
    std::vector<int64_t> plen(np, 0);
    for (int64_t k = 0; k < np; ++k) {
        rc = swb_backtrack(dP + hp_off[k], mm[k] + 1, pos[k], &plen[k], device, nullptr);
        if (rc) die("backtrack", rc);
    }
    t1 = now_s();
    std::printf("Elapsed time for backtracking: %f\n", t1 - t0);
    for (int64_t k = 0; k < np; ++k)
        std::printf("pair %lld: %lld x %lld  maxScore %d  maxPos %lld  path length %lld\n", (long long)k, (long long)mm[k],
                    (long long)nn[k], sc[k], (long long)pos[k], (long long)plen[k]);
    cudaFree(dH); cudaFree(dP); cudaFree(d_pos); cudaFree(d_sc);
    return 0;
}

int main(int argc, char* argv[])
{
    long long m = 8, n = 9;             // omp_smithW.c:70-71
    bool builtin = true, from_files = false;
    std::string a, b;
    const int device = std::getenv("SWB_DEVICE") ? std::atoi(std::getenv("SWB_DEVICE")) : 0;
    if (argc == 3 && std::string(argv[1]) == "--pairs") {
        if (swb_device_count() <= device) { std::fprintf(stderr, "swb: no CUDA device %d (there is no CPU fallback)\n", device); return 1; }
        return run_manifest(argv[2], device);
    }
    if (argc == 4 && std::string(argv[1]) == "--fasta") {
        if (!read_spec(argv[2], a) || !read_spec(argv[3], b)) return 1;
        m = (long long)a.size(); n = (long long)b.size();
        builtin = false; from_files = true;
    } else if (argc == 3) {             // omp_smithW.c:91-96
        m = std::strtoll(argv[1], nullptr, 10);
        n = std::strtoll(argv[2], nullptr, 10);
        builtin = false;
    }
    if (builtin) std::printf("Using built-in data for testing ..\n");
    std::printf("Problem size: Matrix[%lld][%lld], FACTOR=%d CUTOFF=%d\n", n, m, FACTOR, CUTOFF);
    if (m <= 0 || n <= 0) { std::fprintf(stderr, "swb: sizes must be positive\n"); return 1; }
    const bool v1_lines = std::getenv("SWB_V1_LINES") != nullptr;
    const bool skip_bt = std::getenv("SWB_SKIP_BACKTRACK") && std::atoi(std::getenv("SWB_SKIP_BACKTRACK")) != 0;
    if (v1_lines) print_footprint(m, n);

    if (swb_device_count() <= device) { std::fprintf(stderr, "swb: no CUDA device %d (there is no CPU fallback)\n", device); return 1; }

    if (builtin) { a = "TGTTACGG"; b = "GGTTGACTA"; }        // omp_smithW.c:147-164
    else if (!from_files) {
        a.assign((size_t)m, 'A'); b.assign((size_t)n, 'A');
        const unsigned seed = std::getenv("SWB_SEED") ? (unsigned)std::strtoul(std::getenv("SWB_SEED"), nullptr, 10)
                                                      : (unsigned)std::time(nullptr);
        swb_generate(seed, m, n, a.data(), b.data());
    }
    if (skip_bt) {
        // omp_smithW-v1-refinedOrig.cpp:190-192,226-228: no maxPos, no backtrack, nothing to keep -> score-only kernel
        if (v1_lines) std::printf("Skipping backtrack ...\n");
        cudaSetDevice(device); cudaFree(nullptr);
        std::printf("Using %d out of max %d threads...", 1, 1);
        int32_t score = 0; int64_t pos = 0;
        const double s0 = now_s();
        const int rcs = swb_score_only(a.data(), m, b.data(), n, nullptr, &score, &pos, device, nullptr);
        const double s1 = now_s();
        if (rcs) die("score-only fill", rcs);
        std::printf("\nElapsed time for scoring matrix computation: %f\n", s1 - s0);
        std::printf("maxScore: %d  maxPos: %lld\n", score, (long long)pos);
        return 0;
    }
    if (const char* devs = std::getenv("SWB_DEVICES")) {
        // one pair in column strips over several GPUs
        std::vector<int> dl;
        for (const char* p = devs; *p;) { dl.push_back(std::atoi(p)); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
        if (dl.empty() || (long long)dl.size() > m) { std::fprintf(stderr, "swb: bad SWB_DEVICES\n"); return 1; }
        for (int d : dl) if (d < 0 || d >= swb_device_count()) { std::fprintf(stderr, "swb: no CUDA device %d\n", d); return 1; }
        swb_multi* mh = nullptr;
        int rcm = swb_multi_create(&mh, m, n, dl.data(), (int)dl.size());
        if (rcm) die("multi-GPU setup", rcm);
        std::printf("Using %d out of max %d threads...", 1, 1);
        int64_t mp = 0, pl = 0; int32_t ms = 0;
        double m0 = now_s();
        rcm = swb_multi_fill(mh, a.data(), b.data(), nullptr, &mp, &ms);
        double m1 = now_s();
        if (rcm) die("fill", rcm);
        std::printf("\nElapsed time for scoring matrix computation: %f\n", m1 - m0);
        m0 = now_s();
        rcm = swb_multi_backtrack(mh, mp, &pl);
        m1 = now_s();
        if (rcm) die("backtrack", rcm);
        std::printf("Elapsed time for backtracking: %f\n", m1 - m0);
        int exit_code = 0;
        if (builtin) {
            std::vector<int32_t> H((size_t)(m + 1) * (n + 1));
            swb_multi_gather_host(mh, H.data(), nullptr);
            const bool ok = H.back() == 7 && mp == 69;
            std::printf("Verifying results using the builtinIn data: %s\n", ok ? "true" : "false");
            if (!ok) exit_code = 134;
        }
        std::printf("maxPos: %lld  path length: %lld  (%d column strips)\n", (long long)mp, (long long)pl, (int)dl.size());
        swb_multi_destroy(mh);
        return exit_code;
    }

    cudaSetDevice(device);
    const size_t cells = (size_t)(m + 1) * (size_t)(n + 1);
    int32_t *dH = nullptr, *dP = nullptr;
    if (cudaMalloc((void**)&dH, cells * 4) != cudaSuccess || cudaMalloc((void**)&dP, cells * 4) != cudaSuccess) {
        std::fprintf(stderr, "swb: cannot allocate %.2f GB of device memory for H and P\n", cells * 8 / 1e9);
        return 1;
    }
    // warm the context / module so the timer sees the fill, as the reference's timer does (omp_smithW.c:199)
    cudaFree(nullptr);

    std::printf("Using %d out of max %d threads...", 1, 1);       // omp_smithW.c:194 (host threads driving the GPU)
    int64_t maxPos = 0, pathLen = 0;
    double t0 = now_s();
    int rc = swb_fill(a.data(), m, b.data(), n, nullptr, dH, dP, m + 1, &maxPos, device, nullptr);
    double t1 = now_s();
    if (rc) die("fill", rc);
    std::printf("\nElapsed time for scoring matrix computation: %f\n", t1 - t0);   // omp_smithW.c:220

    t0 = now_s();
    rc = swb_backtrack(dP, m + 1, maxPos, &pathLen, device, nullptr);
    t1 = now_s();
    if (rc) die("backtrack", rc);
    std::printf("Elapsed time for backtracking: %f\n", t1 - t0);                    // omp_smithW.c:228

    int exit_code = 0;
    if (builtin) {                                                                   // omp_smithW.c:230-234
        int32_t corner = -1;
        cudaMemcpy(&corner, dH + cells - 1, 4, cudaMemcpyDeviceToHost);
        const bool ok = corner == 7 && maxPos == 69;
        std::printf("Verifying results using the builtinIn data: %s\n", ok ? "true" : "false");
        if (!ok) exit_code = 134;   // the reference aborts on its assert
    }
    if (std::getenv("SWB_DEBUG")) {                                                  // omp_smithW.c:236-242
        std::vector<int32_t> H(cells), P(cells);
        cudaMemcpy(H.data(), dH, cells * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(P.data(), dP, cells * 4, cudaMemcpyDeviceToHost);
        std::printf("\nSimilarity Matrix:\n");
        print_matrix(H, m + 1, n + 1, a, b);
        std::printf("\nPredecessor Matrix:\n");
        print_pred(P, m + 1, n + 1, a, b);
    }
    std::printf("maxPos: %lld  path length: %lld\n", (long long)maxPos, (long long)pathLen);
    cudaFree(dH); cudaFree(dP);
    return exit_code;
}
