/* va_oracle.c -- TEST INFRASTRUCTURE (see va_oracle.h).  CPU restatement of the
 * reference's DP hot path.  Each routine names the reference lines it follows.
 * Written from the behavioural spec in SURVEY.md App. A/B, not from the sources'
 * text: one generic fill routine instead of the reference's eight. */
#include "va_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { P_START = 0, P_UP = 1, P_LEFT = 2, P_DIAG = 3 };

/* Base classes: 0 = anything else (incl. '\0' pad and bytes >= 0x80), 1..4 = A,T,C,G
 * in either case, 5 = N/n.  (DefaultKernel.h:43-60 == opencl_definitions.cl:25-42) */
static int base_class(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 1;
        case 'T': case 't': return 2;
        case 'C': case 'c': return 3;
        case 'G': case 'g': return 4;
        case 'N': case 'n': return 5;
        default: return 0;
    }
}

static int is_acgt(unsigned char c) {
    int k = base_class(c);
    return k >= 1 && k <= 4;
}

/* base_score[6][6] (DefaultKernel.h:83-97): rows/cols 0 and 5 are all zero.  The SIMD
 * kernels get the same numbers from (read&0xDF)==(ref&0xDF) masked by "both are ACGT"
 * (SSEKernel.cpp:341-358). */
int va_oracle_subst(unsigned char a, unsigned char b, const va_oracle_scoring *sc) {
    if (!is_acgt(a) || !is_acgt(b)) return 0;
    return base_class(a) == base_class(b) ? sc->match : sc->mismatch;
}

static inline int16_t wrap16(int v) { return (int16_t)(uint16_t)(unsigned)v; }
static inline int16_t max16(int16_t a, int16_t b) { return a > b ? a : b; }

/* ---- score only ---------------------------------------------------------------- */

/* SW: DefaultKernel.cpp:83-138, SSEKernel.cpp:1007-1150, scoring_kernels.cl:1-96.
 * NW: DefaultKernel.cpp:140-202, SSEKernel.cpp:1152-1315, scoring_kernels.cl:98-198.
 * The stored value is the full short (SSE/AVX/OpenCL); Default only writes its low
 * byte (memset(...,1), DefaultKernel.cpp:137,199). */
static int16_t score_one(int nw, const unsigned char *read, int M, const unsigned char *ref, int N,
                         const va_oracle_scoring *sc, int16_t *rows /* 2*(N+1), zeroed here */) {
    const int W = N + 1;
    memset(rows, 0, sizeof(int16_t) * 2 * (size_t)W);
    int prev = 0, cur = 1;
    int16_t best = 0;
    for (int i = 0; i < M; ++i) {
        int16_t *pr = rows + (size_t)prev * W, *cr = rows + (size_t)cur * W;
        /* column 0 is never written in either mode: stays 0 */
        for (int j = 0; j < N; ++j) {
            int16_t u = wrap16(pr[j + 1] + sc->gap_ref);
            int16_t l = wrap16(cr[j] + sc->gap_read);
            int16_t d = wrap16(pr[j] + va_oracle_subst(read[i], ref[j], sc));
            int16_t h = max16(u, max16(l, d));
            if (!nw) {
                h = max16(h, 0);
                best = max16(best, h);
            }
            cr[j + 1] = h;
        }
        if (nw) best = max16(best, cr[N]); /* last column of every row */
        prev = cur;
        cur ^= 1;
    }
    if (nw) { /* whole last row, column 0 included */
        const int16_t *lr = rows + (size_t)prev * W;
        for (int j = 0; j <= N; ++j) best = max16(best, lr[j]);
    }
    return best;
}

/* ---- fill with pointers + traceback ---------------------------------------------- */

typedef struct {
    int end_read; /* 0-based sequence coordinates of the cell traceback starts from */
    int end_ref;
} end_cell_t;

/* SW fill: DefaultKernel.cpp:204-280 / SSEKernel.cpp:226-453 / alignment_kernels.cl:38-135.
 * NW fill: DefaultKernel.cpp:282-389 / SSEKernel.cpp:455-727 / alignment_kernels.cl:239-364. */
static end_cell_t fill_pointers(int nw, int policy, const unsigned char *read, int M,
                                const unsigned char *ref, int N, const va_oracle_scoring *sc,
                                uint8_t *ptr /* (M+1)*(N+1) */, int16_t *rows) {
    const int W = N + 1;
    memset(rows, 0, sizeof(int16_t) * 2 * (size_t)W);
    memset(ptr, P_START, (size_t)(M + 1) * W);
    int prev = 0, cur = 1;

    /* SW: first strictly greater cell in row-major order */
    int16_t best = 0;
    int best_i = 0, best_j = 0;
    /* NW bookkeeping (App. A.3 steps 3-5) */
    int max_read_pos = M - 1, max_ref_pos = N - 1;
    int row_max_idx = 0, global_row_max_idx = -1;

    for (int i = 0; i < M; ++i) {
        int16_t *pr = rows + (size_t)prev * W, *cr = rows + (size_t)cur * W;
        uint8_t *pp = ptr + (size_t)(i + 1) * W;
        const int read_ok = policy == VA_ORACLE_POLICY_SIMD ? is_acgt(read[i]) : base_class(read[i]) != 0;
        int16_t row_max = 0;
        if (nw) {
            pp[0] = P_UP;
            cr[0] = wrap16((i + 1) * sc->gap_ref);
            if (max_read_pos == M - 1 && !read_ok) max_read_pos = i - 1;
            if (max_read_pos + 1 == i) global_row_max_idx = row_max_idx;
            row_max = cr[0];
            row_max_idx = 0;
        }
        for (int j = 0; j < N; ++j) {
            int16_t u = wrap16(pr[j + 1] + sc->gap_ref);
            int16_t l = wrap16(cr[j] + sc->gap_read);
            int16_t d = wrap16(pr[j] + va_oracle_subst(read[i], ref[j], sc));
            int16_t h = max16(u, max16(l, d));
            if (!nw) h = max16(h, 0);
            cr[j + 1] = h;

            uint8_t p = P_START;
            if (policy == VA_ORACLE_POLICY_DEFAULT_OCL) {
                /* START(if SW and 0) > DIAG > UP > LEFT */
                if (!nw && h == 0) p = P_START;
                else if (h == d) p = P_DIAG;
                else if (h == u) p = P_UP;
                else if (h == l) p = P_LEFT;
            } else {
                /* max of the codes UP=1 < LEFT=2 < DIAG=3; DIAG only between two ACGT bases;
                 * no zero special case */
                if (h == u) p = P_UP;
                if (h == l) p = P_LEFT;
                if (h == d && is_acgt(read[i]) && is_acgt(ref[j])) p = P_DIAG;
            }
            pp[j + 1] = p;

            if (!nw) {
                if (h > best) { best = h; best_i = i; best_j = j; }
            } else {
                const int ref_ok = policy == VA_ORACLE_POLICY_SIMD ? is_acgt(ref[j]) : base_class(ref[j]) != 0;
                if (max_ref_pos == N - 1 && !ref_ok) max_ref_pos = j - 1;
                if (h > row_max) { row_max = h; row_max_idx = j; }
            }
        }
        prev = cur;
        cur ^= 1;
    }
    end_cell_t e;
    if (!nw) {
        e.end_read = best_i;
        e.end_ref = best_j;
    } else {
        if (global_row_max_idx < 0) global_row_max_idx = row_max_idx;
        e.end_read = max_read_pos;
        e.end_ref = max_ref_pos < global_row_max_idx ? max_ref_pos : global_row_max_idx;
    }
    return e;
}

/* DefaultKernel.cpp:391-456,458-525 / SSEKernel.cpp:729-866,868-1005 /
 * alignment_kernels.cl:146-192,370-414.  Returns `start` (== readStart == refStart). */
static int traceback(const uint8_t *ptr, int N, const unsigned char *read, const unsigned char *ref,
                     end_cell_t e, int aln_len, char *out_read, char *out_ref) {
    const int W = N + 1;
    memset(out_read, 0, (size_t)aln_len);
    memset(out_ref, 0, (size_t)aln_len);
    int i = e.end_read, j = e.end_ref, pos = aln_len - 2;
    uint8_t p = ptr[(size_t)(i + 1) * W + (j + 1)];
    while (p != P_START) {
        char a = '-', b = '-';
        if (p == P_UP || p == P_DIAG) a = (char)read[i--];
        if (p == P_LEFT || p == P_DIAG) b = (char)ref[j--];
        if (pos >= 0) { /* the reference would write out of bounds here; never reached with gap scores < 0 */
            out_read[pos] = a;
            out_ref[pos] = b;
        }
        --pos;
        p = ptr[(size_t)(i + 1) * W + (j + 1)];
    }
    return pos + 1;
}

/* ---- batch entry points ------------------------------------------------------------ */

static int pick_threads(int threads) {
#ifdef _OPENMP
    return threads > 0 ? threads : omp_get_max_threads();
#else
    (void)threads;
    return 1;
#endif
}

int va_oracle_score(int opt, int n, const char *reads, int read_length, const char *refs,
                    int ref_length, const va_oracle_scoring *sc, int16_t *scores, int threads) {
    const int alg = opt & 0xF;
    if (alg != VA_ORACLE_SW && alg != VA_ORACLE_NW) return -1;
    const int nt = pick_threads(threads);
#pragma omp parallel num_threads(nt)
    {
        int16_t *rows = (int16_t *)malloc(sizeof(int16_t) * 2 * (size_t)(ref_length + 1));
#pragma omp for schedule(static)
        for (int p = 0; p < n; ++p) {
            scores[p] = score_one(alg == VA_ORACLE_NW,
                                  (const unsigned char *)reads + (size_t)p * read_length, read_length,
                                  (const unsigned char *)refs + (size_t)p * ref_length, ref_length, sc, rows);
        }
        free(rows);
    }
    return 0;
}

int va_oracle_align(int opt, int policy, int n, const char *reads, int read_length,
                    const char *refs, int ref_length, const va_oracle_scoring *sc,
                    char *aln_read, char *aln_ref, int16_t *start, int16_t *end_cell, int threads) {
    const int alg = opt & 0xF;
    if (alg != VA_ORACLE_SW && alg != VA_ORACLE_NW) return -1;
    if (policy != VA_ORACLE_POLICY_DEFAULT_OCL && policy != VA_ORACLE_POLICY_SIMD) return -2;
    const int aln_len = read_length + ref_length;
    const int nt = pick_threads(threads);
#pragma omp parallel num_threads(nt)
    {
        int16_t *rows = (int16_t *)malloc(sizeof(int16_t) * 2 * (size_t)(ref_length + 1));
        uint8_t *ptr = (uint8_t *)malloc((size_t)(read_length + 1) * (size_t)(ref_length + 1));
#pragma omp for schedule(static)
        for (int p = 0; p < n; ++p) {
            const unsigned char *rd = (const unsigned char *)reads + (size_t)p * read_length;
            const unsigned char *rf = (const unsigned char *)refs + (size_t)p * ref_length;
            end_cell_t e = fill_pointers(alg == VA_ORACLE_NW, policy, rd, read_length, rf, ref_length, sc, ptr, rows);
            int st = traceback(ptr, ref_length, rd, rf, e, aln_len, aln_read + (size_t)p * aln_len,
                               aln_ref + (size_t)p * aln_len);
            start[p] = (int16_t)st;
            if (end_cell) {
                end_cell[2 * p] = (int16_t)e.end_read;
                end_cell[2 * p + 1] = (int16_t)e.end_ref;
            }
        }
        free(ptr);
        free(rows);
    }
    return 0;
}
