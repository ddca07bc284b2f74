// va_fast.cu -- the packed inter-task fill kernels: one thread computes TWO pairs at once in the
// two signed 16-bit lanes of every register, with the Blackwell DPX instructions
// (VIADDMNMX.S16x2[.RELU], VIMNMX.S16x2 with predicate outputs, VIMNMX3.S16x2).
//
// Same recurrences as the reference's kernels (score: DefaultKernel.cpp:83-202,
// SSEKernel.cpp:1007-1315; NW fill with pointers: DefaultKernel.cpp:282-389), restructured for
// the GPU instead of translated:
//   * a strip of TW ref columns lives in registers (previous-row H per column + one PRMT selector
//     per column); the thread sweeps all read rows of the strip, then moves to the next strip; the
//     strip's right edge goes through a slot-interleaved boundary array (coalesced 4 B / thread);
//   * the substitution score of both lanes is ONE prmt: the row supplies two 4-byte tables
//     (scores of read base A/B against ref A,C,G,T), the column supplies the selector; the
//     selector's sign-replicate nibbles widen the 8-bit entries to 16-bit lanes;
//   * per cell: t = max(up+gF, left+gR); H = max(diag+s, t [,0])  -> 3 DPX instructions
//     (+ 1 prmt, + 1/2 VIMNMX3 for the SW running maximum);
//   * NW align keeps H+gF per column so that both comparisons the Default/OpenCL pointer rule
//     needs (diag+s >= max(up,left) -> DIAG, else up >= left -> UP, else LEFT) fall out of the
//     two max instructions as predicates; the predicates are banked into bit planes with
//     predicated FADDs (2^23-biased floats: exact integers, runs on the FP pipes and leaves the
//     integer pipe to the recurrence).  2 bits per cell reach HBM, as coalesced 8-byte stores.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "va_fast.cuh"

namespace va {

namespace {

constexpr uint32_t NEG2 = 0x80008000u;  // (-32768, -32768): identity of the packed max

__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { return __viaddmax_s16x2(a, b, NEG2); }

// SYM: gap_read == gap_ref, so "H + gR" (what the cell to the right needs) and "H + gF" (what the
// cell below needs) are the same register: one add less per cell in NW align.
// EDGE (NW align only): the duo's ref is padded (cols < ref_length), so the last true column is kept
// for the traceback kernel's pad-column rule.  It is a separate instantiation -- launched next to the
// plain one, each taking only its own kind of duo -- because the plain kernel sits right at its
// register budget and the extra state costs it spills.
template <int MODE, int TW, bool SYM, bool EDGE>
__global__ void __launch_bounds__(128, 4) fill_fast_kernel(ChunkGeom g, ChunkBuffers b, FastConsts fc) {
    constexpr bool SWA = MODE == MODE_SW_ALIGN;
    constexpr bool NWA = MODE == MODE_NW_ALIGN || SWA;  // both align modes keep H + gF and emit direction planes
    constexpr bool SWS = MODE == MODE_SW_SCORE;
    constexpr bool NWS = MODE == MODE_NW_SCORE;
    constexpr int NG = (TW + 15) / 16;

    __shared__ uint32_t T[8];
    if (threadIdx.x < 8) T[threadIdx.x] = fc.tab[threadIdx.x];
    __syncthreads();

    const int duo = blockIdx.x * blockDim.x + threadIdx.x;
    const int slot_a = 2 * duo, slot_b = slot_a + 1;
    unsigned long long cells = 0;
    bool mine = false;
    PairMeta ma, mb;
    if (slot_b < g.n) {
        ma = b.meta[slot_a];
        mb = b.meta[slot_b];
        mine = duo_is_fast(g, MODE, slot_a, ma, mb);
        if (MODE == MODE_NW_ALIGN) mine = mine && (EDGE == (g.ref_length > ma.cols));
    }
    if (mine) {
        const int m = max((int)ma.rows, (int)mb.rows), n = ma.cols;
        cells = ((unsigned long long)ma.rows + (unsigned long long)mb.rows) * (unsigned long long)n;
        const uint32_t gF2 = fc.gF2, gR2 = fc.gR2, dFR2 = fc.dFR2;
        const uint8_t *cread = reinterpret_cast<const uint8_t *>(b.code_reads);
        const uint8_t *cref = reinterpret_cast<const uint8_t *>(b.code_refs);
        uint32_t *bnd = b.fboundary;
        uint4 *dirs = b.fdirs;

        uint32_t best = 0;  // SW: running max; NW score: max(0, last column, last row)
        // SW align: best cell so far, per lane (first strictly greater in row-major order)
        int gbest_a = 0, gbest_b = 0, gi_a = 0, gi_b = 0, gj_a = 0, gj_b = 0;
        const int nstrips = (n + TW - 1) / TW;
        for (int s = 0; s < nstrips; ++s) {
            const int c0 = s * TW;
            const bool first = s == 0, last = s == nstrips - 1;
            uint32_t sel[TW], H[TW];
#pragma unroll
            for (int k = 0; k < TW; ++k) {
                const int col = min(c0 + k, n - 1);  // columns past n repeat the last one; their cells are never used
                const size_t off = ((size_t)(col >> 4) * g.slots) * 16 + (col & 15);
                const uint32_t fa = cref[off + (size_t)slot_a * 16], fb = cref[off + (size_t)slot_b * 16];
                // nibbles: lane A low byte <- table a[fa], high byte <- its sign; lane B from table b (bytes 4..7)
                sel[k] = fa | ((fa | 8u) << 4) | ((fb | 4u) << 8) | ((fb | 12u) << 12);
                // matrix row 0 is 0 (the align modes keep H + gF; SW align adds its bias, see below)
                H[k] = SWA ? fc.swa_g0 : NWA ? gF2 : 0u;
            }
            // H[i][c0] feeding the first column's diagonal: 0 for matrix row 0
            uint32_t diag_next = SWA ? fc.swa_g0 : NWA ? gF2 : 0u;
            // matrix column 0: 0 in the score modes and SW align, (i+1)*gF in NW align; carried as "left + gR"
            uint32_t col0 = SWA ? fc.swa_l0 : NWA ? add2(gF2, gR2) : gR2;
            // SW align runs the whole recurrence with a constant bias B on every stored value, so that
            //  - the zero floor folds into the two adds:  left = max(h+gR, 0+gR)  (one VIADDMNMX each), and
            //  - "left" is never negative, so  key = left*32 + (31-k)  is one IMAD on the FMA pipe and the
            //    packed maximum of the keys of a row is its best value together with its FIRST column.
            // Per row the key is compared once (value part only) with the strip's best so far.
            int sval_a = fc.swa_off, sval_b = fc.swa_off;  // best (biased) value so far in this strip; swa_off = value 0
            uint32_t skey_a = 0, skey_b = 0;     // key that first exceeded it (gives the column)
            int srow_a = -1, srow_b = -1;        // and its row
            const int kv = min(TW, n - c0);  // valid columns of this strip
            const bool full = kv == TW;

            const uint8_t *ra = cread + (size_t)slot_a * 16;  // read codes of lane A; lane B is the next uint4
            const uint32_t chunk_stride = (uint32_t)g.slots * 16u;
            // One sweep over all read rows of this strip.  PARTIAL (only ever the last strip) guards
            // the few places that must not see the columns past n; full strips run unguarded.
            auto sweep = [&](auto partial_tag) {
                constexpr bool PARTIAL = decltype(partial_tag)::value;
                uint32_t *bp = bnd + duo;
                uint4 *dp = dirs + fast_dir_index(g, s, 0, 0, duo);
                // Software pipeline of the per-row inputs so no row starts by waiting on memory: the
                // read-code bytes are fetched two rows ahead, the row tables (shared-memory look-up by
                // that code) and the boundary word one row ahead.
                auto code_off = [&](int r) { return (uint32_t)(r >> 4) * chunk_stride + (uint32_t)(r & 15); };
                const int mlast = m - 1;
                uint32_t o1 = code_off(min(1, mlast));
                uint32_t ca1 = ra[o1], cb1 = ra[o1 + 16];
                uint32_t o0 = code_off(0);
                uint32_t nta = T[ra[o0]], ntb = T[ra[o0 + 16]];
                uint32_t nleft = first ? 0u : *bp;
                // one matrix row of this strip; NW align returns the row's two direction planes per group
                auto do_row = [&](int i, uint2(&w)[NG]) {
                    const uint32_t ta = nta, tb = ntb;
                    uint32_t left = first ? col0 : nleft;
                    {
                        nta = T[ca1];  // tables of row i+1
                        ntb = T[cb1];
                        const uint32_t o2 = code_off(min(i + 2, mlast));
                        ca1 = ra[o2];  // codes of row i+2
                        cb1 = ra[o2 + 16];
                        if (!first) nleft = bp[i < mlast ? g.duos : 0];  // boundary of row i+1
                    }
                    if (NWA && !SWA && first) col0 = add2(col0, gF2);
                    uint32_t rowkey = 0;
                    uint32_t edge = 0;
                    uint32_t diag = diag_next;
                    diag_next = add2(left, dFR2);
                    float p1l[NG], p1h[NG], p2l[NG], p2h[NG];
                    uint32_t prev_key = 0;
#pragma unroll
                    for (int q = 0; q < NG; ++q) p1l[q] = p1h[q] = p2l[q] = p2h[q] = 8388608.0f;
#pragma unroll
                    for (int k = 0; k < TW; ++k) {
                        const uint32_t sub = prmt(ta, tb, sel[k]);
                        const uint32_t up = H[k];
                        if (NWA) {
                            bool dl, dh, ul, uh;
                            const uint32_t t = __vibmax_s16x2(up, left, &uh, &ul);  // up+gF >= left+gR : UP before LEFT
                            const uint32_t d = add2(diag, sub);                     // diag + s (table holds s - gF)
                            uint32_t h = __vibmax_s16x2(d, t, &dh, &dl);            // diag+s >= max(up,left) : DIAG first
                            const float bit = (float)(1u << (k & 15));
                            if (dl) p1l[k >> 4] += bit;
                            if (dh) p1h[k >> 4] += bit;
                            if (ul) p2l[k >> 4] += bit;
                            if (uh) p2h[k >> 4] += bit;
                            if (SWA) {
                                // the pointer of a positive cell is the NW rule; a zero cell is START, which the
                                // traceback recognises by tracking the score (DefaultKernel.cpp:238-248)
                                left = __viaddmax_s16x2(h, gR2, fc.swa_l0);           // max(h, 0) + gR   (biased)
                                H[k] = SYM ? left : __viaddmax_s16x2(h, gF2, fc.swa_g0);
                                const uint32_t key = left * 32u + (uint32_t)(31 - k) * 0x00010001u;
                                if (PARTIAL) {  // columns past n must not win
                                    if (k < kv) rowkey = __vmaxs2(rowkey, key);
                                } else if (k & 1) {
                                    rowkey = __vimax3_s16x2(rowkey, key, prev_key);
                                } else if (k == TW - 1) {
                                    rowkey = __vmaxs2(rowkey, key);
                                }
                                prev_key = key;
                            } else {
                                left = add2(h, gR2);
                                H[k] = SYM ? left : add2(h, gF2);
                                if (EDGE && PARTIAL && k == kv - 1) edge = left;  // last true column of a partial last strip
                            }
                        } else {
                            const uint32_t t = __viaddmax_s16x2(up, gF2, left);
                            const uint32_t h = SWS ? __viaddmax_s16x2_relu(diag, sub, t) : __viaddmax_s16x2(diag, sub, t);
                            left = add2(h, gR2);
                            H[k] = h;
                            if (SWS) {
                                // running maximum: two cells per VIMNMX3; columns past n (repeats of the
                                // last ref base) stay out of it
                                if (PARTIAL) {
                                    if (k < kv) best = __vmaxs2(best, h);
                                } else if (k & 1) {
                                    best = __vimax3_s16x2(best, h, H[k - 1]);
                                }
                            }
                            if (NWS && PARTIAL && k == kv - 1) best = __vmaxs2(best, h);  // last column of this row
                        }
                        diag = up;
                    }
                    // right edge of the strip for the next strip; NW align also keeps the LAST true column
                    // (the traceback kernel needs it to decide whether the padded arg-max lands in a pad column)
                    if (!last) *bp = left;
                    else if (EDGE) *bp = PARTIAL ? edge : left;
                    bp += g.duos;
                    if (SWA) {
                        // strictly greater VALUE than anything seen before in this strip (rows above): new best cell
                        const int ra_ = (int)((rowkey & 0xFFFFu) >> 5), rb_ = (int)(rowkey >> 21);
                        if (ra_ > sval_a) {
                            sval_a = ra_;
                            srow_a = i;
                            skey_a = rowkey & 0xFFFFu;
                        }
                        if (rb_ > sval_b) {
                            sval_b = rb_;
                            srow_b = i;
                            skey_b = rowkey >> 16;
                        }
                    }
                    if (NWS && !PARTIAL && last) best = __vmaxs2(best, H[TW - 1]);  // last column (SSEKernel.cpp:1285-1291)
                    if (NWA) {
#pragma unroll
                        for (int q = 0; q < NG; ++q) {
                            w[q].x = __byte_perm(__float_as_uint(p1l[q]), __float_as_uint(p1h[q]), 0x5410);
                            w[q].y = __byte_perm(__float_as_uint(p2l[q]), __float_as_uint(p2h[q]), 0x5410);
                        }
                    }
                };
                // two rows per iteration: their direction words leave as one 16-byte store per group
                int i = 0;
                for (; i + 1 < m; i += 2, dp += (size_t)NG * g.duos) {
                    uint2 w0[NG], w1[NG];
                    do_row(i, w0);
                    do_row(i + 1, w1);
                    if (NWA) {
#pragma unroll
                        for (int q = 0; q < NG; ++q) dp[(size_t)q * g.duos] = make_uint4(w0[q].x, w0[q].y, w1[q].x, w1[q].y);
                    }
                }
                if (i < m) {  // odd row count: the last word holds one row
                    uint2 w0[NG];
                    do_row(i, w0);
                    if (NWA) {
#pragma unroll
                        for (int q = 0; q < NG; ++q) dp[(size_t)q * g.duos] = make_uint4(w0[q].x, w0[q].y, 0u, 0u);
                    }
                }
            };
            if (full || (NWA && !SWA && !EDGE)) sweep(std::false_type{});
            else sweep(std::true_type{});
            if (SWA) {
                // fold this strip into the pair's best cell: greater wins; equal wins only from an earlier row
                // (an equal value further right in the same row, or below, comes later in row-major order)
                const int va = sval_a - fc.swa_off, vb = sval_b - fc.swa_off;
                if (srow_a >= 0 && (va > gbest_a || (va == gbest_a && srow_a < gi_a))) {
                    gbest_a = va;
                    gi_a = srow_a;
                    gj_a = c0 + 31 - (int)(skey_a & 31u);
                }
                if (srow_b >= 0 && (vb > gbest_b || (vb == gbest_b && srow_b < gi_b))) {
                    gbest_b = vb;
                    gi_b = srow_b;
                    gj_b = c0 + 31 - (int)(skey_b & 31u);
                }
            }
            if (NWA && !SWA) {  // the row the end-cell rule scans (DefaultKernel.cpp:352-355,381-387)
#pragma unroll
                for (int k = 0; k < TW; ++k)
                    if (c0 + k < n) b.hrow[(size_t)(c0 + k) * g.duos + duo] = H[k];
            }
            if (NWS) {  // whole last row (SSEKernel.cpp:1302-1310); column 0 is 0 and `best` starts at 0
#pragma unroll
                for (int k = 0; k < TW; ++k)
                    if (c0 + k < n) best = __vmaxs2(best, H[k]);
            }
        }
        if (!NWA) {
            b.scores[b.pair_of[slot_a]] = (int16_t)(best & 0xFFFF);
            b.scores[b.pair_of[slot_b]] = (int16_t)(best >> 16);
        }
        if (SWA) {
            const int pa = b.pair_of[slot_a], pb = b.pair_of[slot_b];
            b.scores[pa] = (int16_t)gbest_a;
            b.end_cell[2 * pa] = (int16_t)gi_a;
            b.end_cell[2 * pa + 1] = (int16_t)gj_a;
            b.scores[pb] = (int16_t)gbest_b;
            b.end_cell[2 * pb] = (int16_t)gi_b;
            b.end_cell[2 * pb + 1] = (int16_t)gj_b;
        }
    }
    for (int o = 16; o > 0; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
    if ((threadIdx.x & 31) == 0 && cells) atomicAdd(b.cell_count, cells);
}

template <int MODE, int TW>
void launch_one(const ChunkGeom &g, const ChunkBuffers &b, const FastConsts &fc, cudaStream_t stream) {
    const int threads = 128;
    const int duos = (g.n + 1) / 2;
    const int blocks = (duos + threads - 1) / threads;
    if constexpr (MODE == MODE_NW_ALIGN) {
        // plain duos, then (nothing to do on an unpadded batch) the padded-ref duos
        if (fc.gF == fc.gR) {
            fill_fast_kernel<MODE, TW, true, false><<<blocks, threads, 0, stream>>>(g, b, fc);
            fill_fast_kernel<MODE, TW, true, true><<<blocks, threads, 0, stream>>>(g, b, fc);
        } else {
            fill_fast_kernel<MODE, TW, false, false><<<blocks, threads, 0, stream>>>(g, b, fc);
            fill_fast_kernel<MODE, TW, false, true><<<blocks, threads, 0, stream>>>(g, b, fc);
        }
        return;
    }
    if constexpr (MODE == MODE_SW_ALIGN) {
        if (fc.gF == fc.gR) {
            fill_fast_kernel<MODE, TW, true, false><<<blocks, threads, 0, stream>>>(g, b, fc);
            return;
        }
    }
    fill_fast_kernel<MODE, TW, false, false><<<blocks, threads, 0, stream>>>(g, b, fc);
}

template <int MODE>
void launch_tw(const ChunkGeom &g, const ChunkBuffers &b, const FastConsts &fc, cudaStream_t stream) {
    if constexpr (MODE == MODE_SW_ALIGN) {  // more live state per cell: narrower strips keep it in registers
        switch (g.fast_tw) {
            case 16: launch_one<MODE, 16>(g, b, fc, stream); break;
            default: launch_one<MODE, 20>(g, b, fc, stream); break;
        }
    } else {
        switch (g.fast_tw) {
            case 30: launch_one<MODE, 30>(g, b, fc, stream); break;
            default: launch_one<MODE, 32>(g, b, fc, stream); break;
        }
    }
}

uint32_t table_word(int code, int match, int mismatch, int offset) {
    uint32_t w = 0;
    for (int f = 0; f < 4; ++f) {
        const int s = (code < 4 ? (code == f ? match : mismatch) : 0) - offset;
        w |= ((uint32_t)s & 0xFFu) << (8 * f);
    }
    return w;
}

bool fits8(int v) { return v >= -128 && v <= 127; }

}  // namespace

// The packed kernels are exact only while (a) every table entry fits a signed byte and (b) no
// cell can leave the int16 range; otherwise the call stays on the general 32-bit kernel.
bool fast_scoring_ok(int mode, int policy, const Scoring &sc, int read_length, int ref_length) {
    const bool align = mode == MODE_NW_ALIGN || mode == MODE_SW_ALIGN;
    if (align && policy != 0) return false;  // SSE/AVX pointer rule: general kernel
    int mx = 1;
    for (int v : {sc.match, sc.mismatch, sc.gap_read, sc.gap_ref}) mx = max(mx, v < 0 ? -v : v);
    if (mode == MODE_NW_SCORE || mode == MODE_NW_ALIGN) {
        // shifted recurrence (va_nw.cu): gap scores <= 0, table entries s - gap_ref - gap_read, and
        // 0 <= V <= match*min(rows,cols) + |gap_ref|*rows + |gap_read|*cols
        if (sc.gap_read > 0 || sc.gap_ref > 0) return false;
        const int off = sc.gap_ref + sc.gap_read;
        if (!fits8(sc.match - off) || !fits8(sc.mismatch - off) || !fits8(-off)) return false;
        const long long top = (long long)(sc.match > 0 ? sc.match : 0) * (read_length < ref_length ? read_length : ref_length) -
                              (long long)sc.gap_ref * (read_length + 2) - (long long)sc.gap_read * (ref_length + 2);
        return top + 256 <= 32000;
    }
    const int off = align ? sc.gap_ref : 0;
    if (!fits8(sc.match - off) || !fits8(sc.mismatch - off) || !fits8(-off)) return false;
    // every cell is floored at 0: values stay within [-mx, match * min(rows, cols)]
    const long long top = (long long)(sc.match > 0 ? sc.match : 0) * (read_length < ref_length ? read_length : ref_length);
    // SW align packs (value + bias) * 32 + column into a 16-bit lane: value + 128 + gap must stay below 1024
    if (mode == MODE_SW_ALIGN) return top + 2 * 128 < 1024 && mx <= 100;
    return top + mx <= 32000 && mx <= 8000;
}

int fast_pick_tw(int mode, int ref_length) {
    if (mode == MODE_SW_ALIGN) {
        if (const char *v = getenv("VERSALIGN_CUDA_SWA_TW")) {
            const int t = atoi(v);
            if (t == 16 || t == 20) return t;
        }
        int best_tw = 20, best_cols = 1 << 30;
        for (int t : {20, 16}) {
            const int c = (ref_length + t - 1) / t * t;
            if (c < best_cols) {
                best_cols = c;
                best_tw = t;
            }
        }
        return best_tw;
    }
    // fewest computed columns wins; ties go to the wider strip (fewer passes over the read)
    const int c32 = (ref_length + 31) / 32 * 32, c30 = (ref_length + 29) / 30 * 30;
    return c30 < c32 ? 30 : 32;
}

size_t fast_dirs_bytes_per_row_per_slot(int ref_length) {
    // per row and duo: strips * groups * 8 bytes; both strip widths use 2 groups.  (Rows are stored
    // in pairs; the caller rounds the row count up to even.)
    const size_t s32 = (size_t)(ref_length + 31) / 32, s30 = (size_t)(ref_length + 29) / 30;
    size_t per_duo = (s32 > s30 ? s32 : s30) * 2 * 8;
    // SW align strips: 24 and 20 columns use two groups, 16 columns one
    per_duo = std::max(per_duo, (size_t)((ref_length + 19) / 20) * 2 * 8);
    per_duo = std::max(per_duo, (size_t)((ref_length + 23) / 24) * 2 * 8);
    per_duo = std::max(per_duo, (size_t)((ref_length + 15) / 16) * 1 * 8);
    return per_duo / 2;
}

FastConsts make_fast_consts(int mode, const Scoring &sc) {
    FastConsts fc{};
    const bool align = mode == MODE_NW_ALIGN || mode == MODE_SW_ALIGN;
    const bool nw = mode == MODE_NW_SCORE || mode == MODE_NW_ALIGN;  // shifted recurrence, va_nw.cu
    const int off = nw ? sc.gap_ref + sc.gap_read : align ? sc.gap_ref : 0;
    for (int c = 0; c < 8; ++c) fc.tab[c] = table_word(c, sc.match, sc.mismatch, off);
    fc.gF = sc.gap_ref;
    fc.gR = sc.gap_read;
    fc.gF2 = ((uint32_t)sc.gap_ref & 0xFFFFu) * 0x00010001u;
    fc.gR2 = ((uint32_t)sc.gap_read & 0xFFFFu) * 0x00010001u;
    const int d = align ? sc.gap_ref - sc.gap_read : -sc.gap_read;
    fc.dFR2 = ((uint32_t)d & 0xFFFFu) * 0x00010001u;
    if (mode == MODE_SW_ALIGN) {
        // bias B: every stored value is shifted by it so that "left" = max(h,0)+gR+B is never negative
        const int B = 128;
        auto pk = [](int v) { return ((uint32_t)v & 0xFFFFu) * 0x00010001u; };
        fc.swa_l0 = pk(sc.gap_read + B);   // matrix column 0 / zero floor, as "left"
        fc.swa_g0 = pk(sc.gap_ref + B);    // matrix row 0 / zero floor, as "H + gF"
        fc.swa_off = sc.gap_read + B;      // key>>5 minus this is the cell value
        fc.swa_key0 = pk(((sc.gap_read + B) << 5) | 31);
    }
    return fc;
}

int launch_fill_fast(const ChunkGeom &g, const ChunkBuffers &b, int mode, const Scoring &sc, cudaStream_t stream) {
    if (g.fast_tw == 0 || g.n < 2) return 0;
    const FastConsts fc = make_fast_consts(mode, sc);
    switch (mode) {
        case MODE_SW_SCORE: launch_tw<MODE_SW_SCORE>(g, b, fc, stream); break;
        case MODE_NW_SCORE:
        case MODE_NW_ALIGN: return launch_fill_nw(g, b, mode, fc, stream);
        case MODE_SW_ALIGN: launch_tw<MODE_SW_ALIGN>(g, b, fc, stream); break;
        default: return 0;
    }
    return 1;
}

}  // namespace va
