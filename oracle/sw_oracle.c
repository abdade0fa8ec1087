/*
 * sw_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see sw_oracle.h).
 *
 * Restates omp_smithW.c of chunhualiao/Smith-Waterman; every function cites the
 * reference lines it follows.  Parity: PINNED against the reference's built-in
 * case and against dumps of the unmodified reference (tests/golden/).
 */
#include "sw_oracle.h"

#include <stdlib.h>
#include <string.h>

enum { P_NONE = 0, P_UP = 1, P_LEFT = 2, P_DIAGONAL = 3 }; /* omp_smithW.c:33-36 */

/* omp_smithW.c:495-518: one draw per letter, 0->A 2->C 3->G otherwise T. */
static char draw_letter(void)
{
    switch (rand() % 4) {
    case 0:  return 'A';
    case 2:  return 'C';
    case 3:  return 'G';
    default: return 'T';
    }
}

void swo_generate(unsigned seed, int64_t m, int64_t n, char *a, char *b)
{
    /* omp_smithW.c:491 seeds with time(NULL); the harness pins it. */
    srand(seed);
    /* The reference loops to the already incremented m and n (:109-110), so it
     * consumes m+1 draws for a and n+1 for b.  The extra draw of each is kept
     * in the stream but not stored. */
    for (int64_t k = 0; k <= m; ++k) {
        char c = draw_letter();
        if (k < m) a[k] = c;
    }
    for (int64_t k = 0; k <= n; ++k) {
        char c = draw_letter();
        if (k < n) b[k] = c;
    }
}

int64_t swo_nelement(int64_t i, int64_t M, int64_t N)
{
    /* omp_smithW.c:260-275 */
    int64_t lo = M < N ? M : N;
    int64_t hi = M < N ? N : M;
    if (i < lo) return i;                 /* growing part    */
    if (i < hi) return lo - 1;            /* constant part   */
    return 2 * lo - i + (hi - lo) - 2;    /* shrinking part  */
}

void swo_first_diag_element(int64_t i, int64_t N, int64_t *si, int64_t *sj)
{
    /* omp_smithW.c:282-291: walk down column 1, then along the last row. */
    if (i < N) { *si = i;     *sj = 1; }
    else       { *si = N - 1; *sj = i - N + 2; }
}

/* omp_smithW.c:331-381: one cell.  Order of the strict comparisons is
 * DIAGONAL, UP, LEFT starting from (0, NONE). */
static inline int32_t cell(const char *a, const char *b, const swo_scoring *sc,
                           int64_t M, int64_t i, int64_t j,
                           const int32_t *H, int32_t *pred_out)
{
    int64_t idx = M * i + j;
    int32_t up   = H[idx - M] + sc->gap;
    int32_t left = H[idx - 1] + sc->gap;
    int32_t sub  = (a[j - 1] == b[i - 1]) ? sc->match : sc->mismatch; /* :394-399 */
    int32_t diag = H[idx - M - 1] + sub;
    int32_t best = 0, pred = P_NONE;
    if (diag > best) { best = diag; pred = P_DIAGONAL; }
    if (up   > best) { best = up;   pred = P_UP; }
    if (left > best) { best = left; pred = P_LEFT; }
    *pred_out = pred;
    return best;
}

void swo_fill_wavefront(const char *a, int64_t m, const char *b, int64_t n,
                        const swo_scoring *sc, int32_t *H, int32_t *P,
                        int64_t *maxPos)
{
    const int64_t M = m + 1, N = n + 1;               /* :109-110 */
    memset(H, 0, (size_t)(M * N) * sizeof(int32_t));  /* calloc, :113-118 */
    memset(P, 0, (size_t)(M * N) * sizeof(int32_t));
    int64_t best = 0;                                  /* :173 */
    const int64_t nDiag = M + N - 3;                   /* :182 */
    for (int64_t d = 1; d <= nDiag; ++d) {             /* :203 */
        int64_t cnt = swo_nelement(d, M, N), si, sj;
        swo_first_diag_element(d, N, &si, &sj);
        for (int64_t k = 0; k < cnt; ++k) {            /* :210-215 */
            int64_t i = si - k, j = sj + k;
            int32_t pred;
            int32_t h = cell(a, b, sc, M, i, j, H, &pred);
            H[M * i + j] = h;
            P[M * i + j] = pred;
            if (h > H[best]) best = M * i + j;         /* :384-387 */
        }
    }
    *maxPos = best;
}

void swo_fill_rowmajor(const char *a, int64_t m, const char *b, int64_t n,
                       const swo_scoring *sc, int32_t *H, int32_t *P,
                       int64_t *maxPos)
{
    const int64_t M = m + 1, N = n + 1;
    memset(H, 0, (size_t)M * sizeof(int32_t));
    memset(P, 0, (size_t)M * sizeof(int32_t));
    int32_t gmax = 0;
    int64_t bi = 0, bj = 0;
    for (int64_t i = 1; i < N; ++i) {
        H[M * i] = 0;
        P[M * i] = 0;
        for (int64_t j = 1; j < M; ++j) {
            int32_t pred;
            int32_t h = cell(a, b, sc, M, i, j, H, &pred);
            H[M * i + j] = h;
            P[M * i + j] = pred;
            /* tie-break implied by the wavefront scan: smaller i+j first,
             * then larger i */
            if (h > gmax ||
                (h == gmax && h > 0 &&
                 (i + j < bi + bj || (i + j == bi + bj && i > bi)))) {
                gmax = h; bi = i; bj = j;
            }
        }
    }
    *maxPos = gmax > 0 ? M * bi + bj : 0;
}

void swo_fill_block(const char *a, int64_t m, const char *b, int64_t i0, int64_t i1,
                    const swo_scoring *sc, const int32_t *Htop, int32_t *Hblk, int32_t *Pblk,
                    int32_t *best /* in/out: score, */, int64_t *best_i, int64_t *best_j)
{
    /* rows i0..i1-1 (1-based matrix rows) of the same recurrence as swo_fill_rowmajor,
     * continuing from row i0-1 given in Htop (m+1 ints).  Hblk/Pblk hold (i1-i0) rows. */
    const int64_t M = m + 1;
    int32_t gmax = *best;
    int64_t bi = *best_i, bj = *best_j;
    for (int64_t i = i0; i < i1; ++i) {
        const int32_t *up = (i == i0) ? Htop : Hblk + (i - i0 - 1) * M;
        int32_t *hr = Hblk + (i - i0) * M, *pr = Pblk + (i - i0) * M;
        hr[0] = 0; pr[0] = 0;
        for (int64_t j = 1; j < M; ++j) {
            int32_t sub = (a[j - 1] == b[i - 1]) ? sc->match : sc->mismatch;
            int32_t d = up[j - 1] + sub, u = up[j] + sc->gap, l = hr[j - 1] + sc->gap;
            int32_t h = 0, pred = P_NONE;
            if (d > h) { h = d; pred = P_DIAGONAL; }
            if (u > h) { h = u; pred = P_UP; }
            if (l > h) { h = l; pred = P_LEFT; }
            hr[j] = h; pr[j] = pred;
            if (h > gmax ||
                (h == gmax && h > 0 &&
                 (i + j < bi + bj || (i + j == bi + bj && i > bi)))) {
                gmax = h; bi = i; bj = j;
            }
        }
    }
    *best = gmax; *best_i = bi; *best_j = bj;
}

int64_t swo_backtrack(int32_t *P, int64_t pitch, int64_t maxPos)
{
    /* omp_smithW.c:405-420 */
    if (maxPos <= 0 || P[maxPos] == P_NONE) return 0;
    int64_t len = 0, pos = maxPos;
    do {
        int64_t prev;
        int32_t p = P[pos];
        if (p == P_DIAGONAL)  prev = pos - pitch - 1;
        else if (p == P_UP)   prev = pos - pitch;
        else                  prev = pos - 1;        /* LEFT */
        P[pos] = -p;                                 /* *= PATH */
        ++len;
        pos = prev;
    } while (P[pos] != P_NONE);
    return len;
}

uint64_t swo_fnv1a64(const int32_t *x, int64_t count)
{
    const unsigned char *p = (const unsigned char *)x;
    uint64_t h = 1469598103934665603ull;
    for (int64_t k = 0; k < count * 4; ++k) {
        h ^= p[k];
        h *= 1099511628211ull;
    }
    return h;
}

void swo_score_only(const char *a, int64_t m, const char *b, int64_t n,
                    const swo_scoring *sc, int32_t *maxScore, int64_t *maxPos)
{
    const int64_t M = m + 1;
    int32_t *row = (int32_t *)calloc((size_t)M, sizeof(int32_t));
    int32_t gmax = 0;
    int64_t bi = 0, bj = 0;
    for (int64_t i = 1; i <= n; ++i) {
        int32_t diag_prev = 0;   /* H[i-1][j-1] */
        int32_t left = 0;        /* H[i][j-1]   */
        for (int64_t j = 1; j <= m; ++j) {
            int32_t upv = row[j];
            int32_t sub = (a[j - 1] == b[i - 1]) ? sc->match : sc->mismatch;
            int32_t h = 0, d = diag_prev + sub, u = upv + sc->gap, l = left + sc->gap;
            if (d > h) h = d;
            if (u > h) h = u;
            if (l > h) h = l;
            diag_prev = upv;
            row[j] = h;
            left = h;
            if (h > gmax ||
                (h == gmax && h > 0 &&
                 (i + j < bi + bj || (i + j == bi + bj && i > bi)))) {
                gmax = h; bi = i; bj = j;
            }
        }
    }
    free(row);
    *maxScore = gmax;
    *maxPos = gmax > 0 ? M * bi + bj : 0;
}
