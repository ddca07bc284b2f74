/* va_oracle_affine.c -- TEST INFRASTRUCTURE (see va_oracle.h).  CPU checker for the affine-gap (Gotoh) variant
 * of the hot path.
 *
 * PARITY UNPINNED: the reference has no affine-gap kernel, so there is nothing to execute against.  The variant is
 * defined here as the smallest generalisation of the reference's linear-gap modes (va_oracle.c, pinned): the same
 * borders, end-cell rules and output layout, with a gap of length L costing gap_open + L * gap_{read,ref}.  With
 * gap_open == 0 every score, end cell and alignment equals the linear-gap result (Default/OpenCL pointer policy) --
 * tests/test_affine.py checks that property on the oracle and on the CUDA path, which anchors the variant to the
 * pinned one.
 *
 *   E(i,j) = max(E(i,j-1), H(i,j-1) + gap_open) + gap_read      a gap in the read   (LEFT moves)
 *   F(i,j) = max(F(i-1,j), H(i-1,j) + gap_open) + gap_ref       a gap in the ref    (UP moves)
 *   H(i,j) = max(H(i-1,j-1) + s, F(i,j), E(i,j) [, 0 in SW])
 * Pointers: H comes from START (SW and 0) > DIAG > F > E  (the linear rule's DIAG > UP > LEFT); a gap state opens
 * (returns to H) when opening is at least as good as extending.  Arithmetic is plain int; the exactness domain is the
 * library's (no H leaves the int16 range). */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "va_oracle.h"

#define NEG (-(1 << 29))
enum { H_START = 0, H_UP = 1, H_LEFT = 2, H_DIAG = 3, E_OPEN = 4, F_OPEN = 8 };

static int imax(int a, int b) { return a > b ? a : b; }
static int valid_default(unsigned char c) {  /* char_to_score != 0: ACGT and N, either case (DefaultKernel.h:43-60) */
    switch (c) {
        case 'A': case 'a': case 'C': case 'c': case 'G': case 'g': case 'T': case 't': case 'N': case 'n': return 1;
        default: return 0;
    }
}

typedef struct { int score, end_read, end_ref; } result_t;

/* one pair; ptr == NULL: score only */
static result_t fill(int nw, int align, const unsigned char *read, int M, const unsigned char *ref, int N,
                     const va_oracle_scoring *sc, int gO, uint8_t *ptr, int *Hrow, int *Frow) {
    const int W = N + 1, gR = sc->gap_read, gF = sc->gap_ref;
    for (int j = 0; j <= N; ++j) { Hrow[j] = 0; Frow[j] = NEG; }
    if (ptr) memset(ptr, H_START, (size_t)(M + 1) * W);
    int best = 0, best_i = 0, best_j = 0;                 /* SW */
    int border = 0;                                        /* NW score */
    int max_read_pos = M - 1, max_ref_pos = N - 1, row_max_idx = 0, global_row_max_idx = -1, row_max = 0, last_row_max = 0;
    for (int i = 0; i < M; ++i) {
        int diag = Hrow[0];
        int left = 0, e = NEG;
        if (nw && align) {
            left = gO + (i + 1) * gF;  /* column 0: a leading gap in the ref, UP pointers (DefaultKernel.cpp:304) */
            if (ptr) ptr[(size_t)(i + 1) * W] = H_UP | F_OPEN * (i == 0);
            if (max_read_pos == M - 1 && !valid_default(read[i])) max_read_pos = i - 1;
            if (max_read_pos + 1 == i) { global_row_max_idx = row_max_idx; last_row_max = row_max; }
            row_max = left;
            row_max_idx = 0;
        }
        Hrow[0] = left;
        for (int j = 0; j < N; ++j) {
            const int up = Hrow[j + 1];
            const int eo = left + gO >= e, fo = up + gO >= Frow[j + 1];
            e = imax(e, left + gO) + gR;
            const int f = imax(Frow[j + 1], up + gO) + gF;
            const int d = diag + va_oracle_subst(read[i], ref[j], sc);
            int h = imax(d, imax(f, e));
            if (!nw) h = imax(h, 0);
            if (ptr) {
                uint8_t p = h == d ? H_DIAG : (h == f ? H_UP : H_LEFT);
                if (!nw && h == 0) p = H_START;
                ptr[(size_t)(i + 1) * W + j + 1] = (uint8_t)(p | (eo ? E_OPEN : 0) | (fo ? F_OPEN : 0));
            }
            if (!nw) {
                if (h > best) { best = h; best_i = i; best_j = j; }
            } else if (!align) {
                if (j == N - 1 || i == M - 1) border = imax(border, h);
            } else {
                if (max_ref_pos == N - 1 && !valid_default(ref[j])) max_ref_pos = j - 1;
                if (h > row_max) { row_max = h; row_max_idx = j; }
            }
            diag = up;
            Hrow[j + 1] = h;
            Frow[j + 1] = f;
            left = h;
        }
    }
    result_t r;
    if (!nw) {
        r.score = best; r.end_read = best_i; r.end_ref = best_j;
    } else if (!align) {
        r.score = border; r.end_read = r.end_ref = 0;
    } else {
        if (global_row_max_idx < 0) { global_row_max_idx = row_max_idx; last_row_max = row_max; }
        r.score = last_row_max;
        r.end_read = max_read_pos;
        r.end_ref = max_ref_pos < global_row_max_idx ? max_ref_pos : global_row_max_idx;
    }
    return r;
}

static int traceback(int nw, const uint8_t *ptr, int N, const unsigned char *read, const unsigned char *ref, result_t e,
                     int aln_len, char *out_read, char *out_ref) {
    const int W = N + 1;
    memset(out_read, 0, (size_t)aln_len);
    memset(out_ref, 0, (size_t)aln_len);
    int i = e.end_read, j = e.end_ref, pos = aln_len - 2, state = 0; /* 0 = H, 1 = F (UP run), 2 = E (LEFT run) */
    (void)nw;
    for (;;) {
        const uint8_t p = ptr[(size_t)(i + 1) * W + (j + 1)];
        char a = '-', b = '-';
        if (state == 0) {
            const int hp = p & 3;
            if (hp == H_START) break;
            if (hp == H_UP) { state = 1; continue; }
            if (hp == H_LEFT) { state = 2; continue; }
            a = (char)read[i--];
            b = (char)ref[j--];
        } else if (state == 1) {
            a = (char)read[i];
            if (p & F_OPEN) state = 0;
            --i;
        } else {
            b = (char)ref[j];
            if (p & E_OPEN) state = 0;
            --j;
        }
        if (pos >= 0) { out_read[pos] = a; out_ref[pos] = b; }
        --pos;
    }
    return pos + 1;
}

static int threads_of(int threads) {
#ifdef _OPENMP
    return threads > 0 ? threads : omp_get_max_threads();
#else
    (void)threads;
    return 1;
#endif
}

int va_oracle_score_affine(int opt, int n, const char *reads, int read_length, const char *refs, int ref_length,
                           const va_oracle_scoring *sc, int gap_open, int16_t *scores, int threads) {
    const int alg = opt & 0xF;
    if (alg != VA_ORACLE_SW && alg != VA_ORACLE_NW) return -1;
#pragma omp parallel num_threads(threads_of(threads))
    {
        int *Hrow = (int *)malloc(sizeof(int) * 2 * (size_t)(ref_length + 1)), *Frow = Hrow + ref_length + 1;
#pragma omp for schedule(static)
        for (int p = 0; p < n; ++p)
            scores[p] = (int16_t)fill(alg == VA_ORACLE_NW, 0, (const unsigned char *)reads + (size_t)p * read_length, read_length,
                                      (const unsigned char *)refs + (size_t)p * ref_length, ref_length, sc, gap_open, NULL, Hrow, Frow).score;
        free(Hrow);
    }
    return 0;
}

int va_oracle_align_affine(int opt, int n, const char *reads, int read_length, const char *refs, int ref_length,
                           const va_oracle_scoring *sc, int gap_open, char *aln_read, char *aln_ref, int16_t *start,
                           int16_t *end_cell, int16_t *scores, int threads) {
    const int alg = opt & 0xF;
    if (alg != VA_ORACLE_SW && alg != VA_ORACLE_NW) return -1;
    const int aln_len = read_length + ref_length;
#pragma omp parallel num_threads(threads_of(threads))
    {
        int *Hrow = (int *)malloc(sizeof(int) * 2 * (size_t)(ref_length + 1)), *Frow = Hrow + ref_length + 1;
        uint8_t *ptr = (uint8_t *)malloc((size_t)(read_length + 1) * (size_t)(ref_length + 1));
#pragma omp for schedule(static)
        for (int p = 0; p < n; ++p) {
            const unsigned char *rd = (const unsigned char *)reads + (size_t)p * read_length;
            const unsigned char *rf = (const unsigned char *)refs + (size_t)p * ref_length;
            const result_t e = fill(alg == VA_ORACLE_NW, 1, rd, read_length, rf, ref_length, sc, gap_open, ptr, Hrow, Frow);
            start[p] = (int16_t)traceback(alg == VA_ORACLE_NW, ptr, ref_length, rd, rf, e, aln_len, aln_read + (size_t)p * aln_len,
                                          aln_ref + (size_t)p * aln_len);
            if (end_cell) { end_cell[2 * p] = (int16_t)e.end_read; end_cell[2 * p + 1] = (int16_t)e.end_ref; }
            if (scores) scores[p] = (int16_t)e.score;
        }
        free(ptr);
        free(Hrow);
    }
    return 0;
}
