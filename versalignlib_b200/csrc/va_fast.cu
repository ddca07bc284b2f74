// va_fast.cu -- the packed inter-task fill kernels of the Smith-Waterman modes: one thread computes
// TWO pairs at once in the two signed 16-bit lanes of every register, with the Blackwell DPX
// instructions (VIADDMNMX.S16x2[.RELU], VIMNMX.S16x2 with predicate outputs, VIMNMX3.S16x2).
// (The NW modes use a shifted recurrence with fewer instructions per cell: va_nw.cu.)
//
// Same recurrences as the reference's kernels (score: DefaultKernel.cpp:83-138,
// SSEKernel.cpp:1007-1150; fill with pointers: DefaultKernel.cpp:204-280), restructured for the GPU
// instead of translated:
//   * a strip of TW ref columns lives in registers (previous-row H per column + one PRMT selector
//     per column); the thread sweeps all read rows of the strip, then moves to the next strip; the
//     strip's right edge goes through a slot-interleaved boundary array (coalesced 4 B / thread);
//   * the substitution score of both lanes is ONE prmt: the row supplies two 4-byte tables
//     (scores of read base A/B against ref A,C,G,T), the column supplies the selector; the
//     selector's sign-replicate nibbles widen the 8-bit entries to 16-bit lanes;
//   * per cell: t = max(up+gF, left+gR); H = max(diag+s, t, 0)  -> 3 DPX instructions
//     (+ 1 prmt, + 1/2 VIMNMX3 for the running maximum);
//   * SW align keeps H+gF per column so that both comparisons the Default/OpenCL pointer rule
//     needs (diag+s >= max(up,left) -> DIAG, else up >= left -> UP, else LEFT) fall out of the
//     two max instructions as predicates; the predicates are banked into bit planes with
//     predicated FADDs (2^23-biased floats: exact integers, runs on the FP pipes and leaves the
//     integer pipe to the recurrence).  2 bits per cell reach HBM, as coalesced 16-byte stores;
//   * the per-row inputs are staged 16 rows ahead with cp.async (see `sweep`).
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "va_fast.cuh"

namespace va {

namespace {

constexpr uint32_t NEG2 = 0x80008000u;  // (-32768, -32768): identity of the packed max

__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { return __viaddmax_s16x2(a, b, NEG2); }

// Ampere-style asynchronous global -> shared copies (LDGSTS): the data never passes through a register.
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Threads per block and register cap per instantiation (4 x 128 threads per SM leave 128 registers each,
// 5 x 96 leave 136): see va_nw.cu -- the pairs whose row loops ptxas compiles without parking predicates
// (tools/check_sass.py).
#ifndef VA_FAST_FORCE_NT
template <int MODE, int TW, bool SYM, bool SOLO, int POLICY = 0>
struct FastBlock {  // SW align duo kernel with equal gaps: 4 x 128 threads, 128 registers; everything else 5 x 96, 136
                    // (score: measured 1-3 % faster; SW align's solo / asymmetric-gap loops: clean only there)
    static constexpr bool WIDE = MODE == MODE_SW_SCORE || (MODE == MODE_SW_ALIGN && (SOLO || !SYM || POLICY == 1));
    static constexpr int NT = WIDE ? 96 : 128;
    static constexpr int MAXREG = WIDE ? 136 : 128;
};
#else
template <int MODE, int TW, bool SYM, bool SOLO, int POLICY = 0>
struct FastBlock {
    static constexpr int NT = VA_FAST_FORCE_NT;
    static constexpr int MAXREG = VA_FAST_FORCE_NT == 96 ? 136 : 128;
};
#endif

// The Smith-Waterman kernels (the NW modes run on a shifted recurrence, va_nw.cu).
// SYM: gap_read == gap_ref, so "H + gR" (what the cell to the right needs) and "H + gF" (what the
// cell below needs) are the same register: one add less per cell in SW align.
// SOLO: the instantiation for single slots whose duo is not fast (va_fast.cuh): same sweep, the owner's
// halves of the shared words stored with 16-bit stores.
// POLICY (SW align): 0 = Default / OpenCL pointers (DIAG > UP > LEFT; a zero cell is START, recognised by the traceback
// from the score along the path).  1 = SSE / AVX (SSEKernel.cpp:366-379: DIAG > LEFT > UP, no zero rule -- a cell is
// START only when its value 0 beats all three candidates, so the walk runs on through zero cells): the second plane
// records LEFT >= UP and a THIRD plane records "h >= 0" (not START), one more max with predicate outputs per cell.
template <int MODE, int TW, bool SYM, bool SOLO, int POLICY = 0>
__global__ void __maxnreg__((FastBlock<MODE, TW, SYM, SOLO, POLICY>::MAXREG)) fill_fast_kernel(ChunkGeom g, ChunkBuffers b, FastConsts fc) {
    constexpr bool SWA = MODE == MODE_SW_ALIGN;
    constexpr int NG = (TW + 15) / 16;
    static_assert(MODE == MODE_SW_ALIGN || MODE == MODE_SW_SCORE, "SW modes only");

    __shared__ uint2 s_T2[256];              // [7*code_a + code_b] -> the two lanes' 4-entry score tables (49 used)
    constexpr int NT = FastBlock<MODE, TW, SYM, SOLO, POLICY>::NT;
    constexpr bool ZPLANE = SWA && POLICY == 1;
    __shared__ uint4 s_idx[3][NT];           // staged row indices: 16 rows per thread and buffer
    __shared__ uint32_t s_bnd[3][16][NT];    // staged right edge of the previous strip, [row][thread]
    for (int t = threadIdx.x; t < 256; t += NT) s_T2[t] = t < 49 ? make_uint2(fc.tab[t / 7], fc.tab[t % 7]) : make_uint2(0u, 0u);
    __syncthreads();

    unsigned long long cells = 0;
    // one work item: a duo (both lanes) or, in the SOLO instantiation, one slot (one lane)
    auto run = [&](const FastWork &fw) {
        const int duo = fw.duo;
        const int slot_a = 2 * duo, slot_b = slot_a + 1;
        const int m = fw.rows, n = fw.cols;
        cells += !SOLO ? ((unsigned long long)fw.ma.rows + (unsigned long long)fw.mb.rows) * (unsigned long long)n
                      : (unsigned long long)m * (unsigned long long)n;
        const uint32_t gF2 = fc.gF2, gR2 = fc.gR2, dFR2 = fc.dFR2, k32 = fc.swa_k32;
        const uint8_t *cref = reinterpret_cast<const uint8_t *>(b.code_refs);
        uint32_t *bnd = b.fboundary;
        uint4 *dirs = b.fdirs;
        uint2 *zdirs = b.fdirs_z;

        uint32_t best = 0;  // SW score: running max
        // SW align: best cell so far, per lane (first strictly greater in row-major order)
        int gbest_a = 0, gbest_b = 0, gi_a = 0, gi_b = 0, gj_a = 0, gj_b = 0;
        const int nstrips = (n + TW - 1) / TW;
        for (int s = 0; s < nstrips; ++s) {
            const int c0 = s * TW;
            const bool first = s == 0, last = s == nstrips - 1;
            uint32_t sel[TW], H[TW];
            // (score mode only: the align kernels' row loops sit at a register-allocation cliff, their prologue stays as it is)
            constexpr bool LEAN = !SWA;
            if (LEAN && c0 + TW <= n) {  // full strip, the common case: va_fast.cuh
                fast_strip_selectors<TW>(g, b.code_refs, slot_a, c0, sel);
#pragma unroll
                for (int k = 0; k < TW; ++k) H[k] = 0u;  // matrix row 0
            } else {  // partial strip (and the kernels that keep the byte-wise prologue)
#pragma unroll
                for (int k = 0; k < TW; ++k) {
                    // Columns past n (partial last strip) get a selector that yields s <= 0 for both lanes (the sign
                    // byte of a table entry): with both gap scores <= 0 such a cell never exceeds the real cells
                    // above / left of it, so the running maximum and the best-cell search need no per-column guards.
                    const int col = min(c0 + k, n - 1);
                    const size_t off = ((size_t)(col >> 4) * g.slots) * 16 + (col & 15);
                    const uint32_t fa = cref[off + (size_t)slot_a * 16], fb = cref[off + (size_t)slot_b * 16];
                    // nibbles: lane A low byte <- table a[fa], high byte <- its sign; lane B from table b (bytes 4..7)
                    sel[k] = c0 + k < n ? (fa | ((fa | 8u) << 4) | ((fb | 4u) << 8) | ((fb | 12u) << 12)) : 0xCC88u;
                    // matrix row 0 is 0 (SW align keeps H + gF, biased, see below)
                    H[k] = SWA ? fc.swa_g0 : 0u;
                }
            }
            // H[i][c0] feeding the first column's diagonal: 0 for matrix row 0
            uint32_t diag_next = SWA ? fc.swa_g0 : 0u;
            // matrix column 0 is 0; carried as "left + gR"
            const uint32_t col0 = SWA ? fc.swa_l0 : gR2;
            // SW align runs the whole recurrence with a constant bias B on every stored value, so that
            //  - the zero floor folds into the two adds:  left = max(h+gR, 0+gR)  (one VIADDMNMX each), and
            //  - "left" is never negative, so  key = left*32 + (31-k)  is one IMAD on the FMA pipe and the
            //    packed maximum of the keys of a row is its best value together with its FIRST column.
            // Per row the key is compared once (value part only) with the strip's best so far.
            int sval_a = fc.swa_off, sval_b = fc.swa_off;  // best (biased) value so far in this strip; swa_off = value 0
            uint32_t skey_a = 0, skey_b = 0;     // key that first exceeded it (gives the column)
            int srow_a = -1, srow_b = -1;        // and its row
            // One sweep over all read rows of this strip.
            //
            // Per-row inputs (the row's pair of substitution tables, the previous strip's right edge) never
            // make a row wait on HBM: they are staged 16 rows ahead with cp.async into this thread's own
            // shared-memory slots -- no registers held across the loop, nothing for the scheduler to sink
            // next to the consumer -- one commit group per 16-row chunk, three buffers.
            {
                uint32_t *bp = bnd + duo;
                uint4 *dp = dirs + fast_dir_index(g, s, 0, 0, duo);
                uint2 *zp = ZPLANE ? zdirs + fast_dir_index(g, s, 0, 0, duo) : nullptr;
                const int nchunks = (m + 15) >> 4;
                auto stage = [&](int c, int buf) {
                    cp_async16(&s_idx[buf][threadIdx.x], b.row_idx + (size_t)c * g.duos + duo);
                    if (!first) {
                        const uint32_t *src = bnd + (size_t)(c * 16) * g.duos + duo;
#pragma unroll
                        for (int r = 0; r < 16; ++r)
                            // (score mode: no row test -- the boundary block is allocated 16 rows past the last one)
                            if (LEAN || c * 16 + r < m) cp_async4(&s_bnd[buf][r][threadIdx.x], src + (size_t)r * g.duos);
                    }
                    cp_async_commit();
                };
                // the third plane of one row (32 bits per group: 16 columns x both lanes) into its half of the row pair's word
                auto store_z = [&](int half, const uint32_t(&z)[NG]) {
#pragma unroll
                    for (int q = 0; q < NG; ++q) {
                        uint32_t *zw = reinterpret_cast<uint32_t *>(zp + (size_t)q * g.duos);
                        if (!SOLO) zw[half] = z[q];
                        else reinterpret_cast<uint16_t *>(zw)[2 * half + fw.lane] = (uint16_t)(z[q] >> (16 * fw.lane));
                    }
                };
                // one matrix row of this strip; SW align returns the row's two direction planes per group (+ the third in wz)
                uint32_t wz[NG];
                auto do_row = [&](int i, const uint2 tt, uint32_t left_in, uint2(&w)[NG]) {
                    const uint32_t ta = tt.x, tb = tt.y;
                    // A solo thread shares its boundary words with another thread: the half that is not its own
                    // holds whatever was there.  SW align's key (left * 32) is the one operation that crosses
                    // lanes, so that half is replaced by a sane value before it enters the recurrence.
                    if (SOLO && SWA) left_in = fw.lane ? ((left_in & 0xFFFF0000u) | (col0 & 0x0000FFFFu)) : ((left_in & 0x0000FFFFu) | (col0 & 0xFFFF0000u));
                    uint32_t left = first ? col0 : left_in;
                    uint32_t rowkey = 0;
                    uint32_t diag = diag_next;
                    diag_next = add2(left, dFR2);
                    float p1l[NG], p1h[NG], p2l[NG], p2h[NG], p3l[NG], p3h[NG];
                    uint32_t prev_key = 0;
#pragma unroll
                    for (int q = 0; q < NG; ++q) p1l[q] = p1h[q] = p2l[q] = p2h[q] = p3l[q] = p3h[q] = 8388608.0f;
#pragma unroll
                    for (int k = 0; k < TW; ++k) {
                        const uint32_t sub = prmt(ta, tb, sel[k]);
                        const uint32_t up = H[k];
                        if (SWA) {
                            bool dl, dh, ul, uh;
                            // policy 0: up+gF >= left+gR -> UP before LEFT; policy 1: the other way round
                            const uint32_t t = POLICY == 0 ? __vibmax_s16x2(up, left, &uh, &ul) : __vibmax_s16x2(left, up, &uh, &ul);
                            const uint32_t d = add2(diag, sub);                     // diag + s (table holds s - gF)
                            uint32_t h = __vibmax_s16x2(d, t, &dh, &dl);            // diag+s >= max(up,left) : DIAG first
                            const float bit = (float)(1u << (k & 15));
                            if (dl) p1l[k >> 4] += bit;
                            if (dh) p1h[k >> 4] += bit;
                            if (ul) p2l[k >> 4] += bit;
                            if (uh) p2h[k >> 4] += bit;
                            // the pointer of a positive cell is the NW rule; a zero cell is START, which the
                            // traceback recognises by tracking the score (DefaultKernel.cpp:238-248)
                            if (ZPLANE) {
                                bool zh, zl;
                                h = __vibmax_s16x2(h, fc.swa_zero, &zh, &zl);      // h >= 0: the cell has a pointer (not START)
                                if (zl) p3l[k >> 4] += bit;
                                if (zh) p3h[k >> 4] += bit;
                                left = add2(h, gR2);
                                H[k] = SYM ? left : add2(h, gF2);
                            } else {
                                left = __viaddmax_s16x2(h, gR2, fc.swa_l0);       // max(h, 0) + gR   (biased)
                                H[k] = SYM ? left : __viaddmax_s16x2(h, gF2, fc.swa_g0);
                            }
                            const uint32_t key = left * k32 + (uint32_t)(31 - k) * 0x00010001u;  // k32 = 32, opaque: stays an IMAD (FMA pipe)
                            if (k & 1) {
                                rowkey = __vimax3_s16x2(rowkey, key, prev_key);
                            } else if (k == TW - 1) {
                                rowkey = __vmaxs2(rowkey, key);
                            }
                            prev_key = key;
                        } else {
                            const uint32_t t = __viaddmax_s16x2(up, gF2, left);
                            const uint32_t h = __viaddmax_s16x2_relu(diag, sub, t);
                            left = add2(h, gR2);
                            H[k] = h;
                            // running maximum: two cells per VIMNMX3
                            if (k & 1) best = __vimax3_s16x2(best, h, H[k - 1]);
                        }
                        diag = up;
                    }
                    // right edge of the strip for the next strip
                    if (!last) store_lanes<SOLO>(bp, left, fw);
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
#pragma unroll
                        for (int q = 0; q < NG; ++q) {
                            w[q].x = __byte_perm(__float_as_uint(p1l[q]), __float_as_uint(p1h[q]), 0x5410);
                            w[q].y = __byte_perm(__float_as_uint(p2l[q]), __float_as_uint(p2h[q]), 0x5410);
                            if (ZPLANE) wz[q] = __byte_perm(__float_as_uint(p3l[q]), __float_as_uint(p3h[q]), 0x5410);
                        }
                    }
                };
                // Software pipeline on top of the staging: the tables and the left edge of the two rows of an
                // iteration are fetched from shared memory during the iteration before, so chunks c and c+1
                // must both have landed while chunk c is swept (three buffers, chunk c+2 in flight).
                stage(0, 0);
                if (nchunks > 1) stage(1, 1);
                else cp_async_commit();
                cp_async_wait<1>();  // chunk 0
                const uint8_t *ib = reinterpret_cast<const uint8_t *>(&s_idx[0][threadIdx.x]);
                const uint32_t *lb = &s_bnd[0][0][threadIdx.x];
                uint2 nt0 = s_T2[ib[0]], nt1 = s_T2[ib[1]];
                uint32_t nl0 = lb[0], nl1 = lb[NT];
                int buf = 0;
                for (int c = 0; c < nchunks; ++c) {
                    const int buf1 = buf == 2 ? 0 : buf + 1, buf2 = buf1 == 2 ? 0 : buf1 + 1;
                    if (c + 2 < nchunks) stage(c + 2, buf2);
                    else cp_async_commit();
                    cp_async_wait<1>();  // chunk c+1 has landed; only the chunk just requested may be in flight
                    const uint8_t *ip = reinterpret_cast<const uint8_t *>(&s_idx[buf][threadIdx.x]) + 2;
                    const uint32_t *lp = &s_bnd[buf][2][threadIdx.x];
                    const uint8_t *ip_next = reinterpret_cast<const uint8_t *>(&s_idx[buf1][threadIdx.x]);
                    const uint32_t *lp_next = &s_bnd[buf1][0][threadIdx.x];
                    const int r0 = c * 16, rend = min(16, m - r0);
                    // two rows per iteration: their direction words leave as one 16-byte store per group
                    int r = 0;
                    for (; r + 1 < rend; r += 2, dp += (size_t)NG * g.duos, zp += ZPLANE ? (size_t)NG * g.duos : 0) {
                        const uint2 t0 = nt0, t1 = nt1;
                        const uint32_t l0 = nl0, l1 = nl1;
                        {  // rows r+2, r+3 (the first two rows of the next chunk after rows 14, 15)
                            const uint8_t *pi = r == 14 ? ip_next : ip;
                            const uint32_t *pl = r == 14 ? lp_next : lp;
                            nt0 = s_T2[pi[0]];
                            nt1 = s_T2[pi[1]];
                            nl0 = pl[0];
                            nl1 = pl[NT];
                            ip += 2;
                            lp += 2 * NT;
                        }
                        // each row's planes leave as soon as the row is done (8-byte halves of the row pair's
                        // 16-byte word; L2 merges them), so no plane register lives across the other row
                        uint2 w0[NG];
                        do_row(r0 + r, t0, l0, w0);
                        if (SWA) {
#pragma unroll
                            for (int q = 0; q < NG; ++q) store_half<SOLO>(dp + (size_t)q * g.duos, 0, w0[q], fw);
                            if (ZPLANE) store_z(0, wz);
                        }
                        do_row(r0 + r + 1, t1, l1, w0);
                        if (SWA) {
#pragma unroll
                            for (int q = 0; q < NG; ++q) store_half<SOLO>(dp + (size_t)q * g.duos, 1, w0[q], fw);
                            if (ZPLANE) store_z(1, wz);
                        }
                    }
                    if (r < rend) {  // odd row count (last chunk only): the last word holds one row
                        uint2 w0[NG];
                        do_row(r0 + r, nt0, nl0, w0);
                        if (SWA) {
#pragma unroll
                            for (int q = 0; q < NG; ++q) store_half<SOLO>(dp + (size_t)q * g.duos, 0, w0[q], fw);
                            if (ZPLANE) store_z(0, wz);
                        }
                    }
                    buf = buf1;
                }
                cp_async_wait<0>();
            }
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
        }
        const bool out_a = !SOLO || fw.lane == 0, out_b = !SOLO || fw.lane == 1;
        if (!SWA) {
            if (out_a) b.scores[b.pair_of[slot_a]] = (int16_t)(best & 0xFFFF);
            if (out_b) b.scores[b.pair_of[slot_b]] = (int16_t)(best >> 16);
        } else {
            if (out_a) {
                const int pa = b.pair_of[slot_a];
                b.scores[pa] = (int16_t)gbest_a;
                b.end_cell[2 * pa] = (int16_t)gi_a;
                b.end_cell[2 * pa + 1] = (int16_t)gj_a;
            }
            if (out_b) {
                const int pb = b.pair_of[slot_b];
                b.scores[pb] = (int16_t)gbest_b;
                b.end_cell[2 * pb] = (int16_t)gi_b;
                b.end_cell[2 * pb + 1] = (int16_t)gj_b;
            }
        }
    };
    const int thread = blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (!SOLO) {  // thread t takes duo t
        // slots are sorted by ascending extents: blocks are taken from the far end so the longest pairs start
        // first and the last wave is made of the short ones
        const FastWork fw = fast_work_duo(g, b.meta, MODE, (int)(gridDim.x - 1 - blockIdx.x) * (int)blockDim.x + (int)threadIdx.x);
        if (fw.own == OWN_DUO) run(fw);
    } else {  // grid-stride loop over the slots the prep kernel listed
        const int count = *b.solo_count;
        for (int e = thread; e < count; e += (int)(gridDim.x * blockDim.x)) run(fast_work_solo(b.meta, b.solo_list[e]));
    }
    for (int o = 16; o > 0; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
    if ((threadIdx.x & 31) == 0 && cells) atomicAdd(b.cell_count, cells);
}

template <int MODE, int TW, bool SYM, bool SOLO, int POLICY = 0>
void launch_inst(const ChunkGeom &g, const ChunkBuffers &b, const FastConsts &fc, cudaStream_t stream) {
    const int threads = FastBlock<MODE, TW, SYM, SOLO, POLICY>::NT;
    const int duos = (g.n + 1) / 2;
    const int blocks = (duos + threads - 1) / threads;
    // duo grid: one thread per duo.  solo grid: a fixed grid strides over the list the prep kernel compiled
    // (va_fast.cuh); empty on a uniform batch, where the blocks read the count and leave.
    fill_fast_kernel<MODE, TW, SYM, SOLO, POLICY><<<SOLO ? std::min(2 * blocks, 148 * 4) : blocks, threads, 0, stream>>>(g, b, fc);
}

template <int MODE, int TW>
void launch_one(const ChunkGeom &g, const ChunkBuffers &b, const FastConsts &fc, cudaStream_t stream) {
    // SW align with equal gap scores keeps one register per column less
    if (MODE == MODE_SW_ALIGN && fc.gF == fc.gR) {
        if (g.n >= 2) launch_inst<MODE, TW, MODE == MODE_SW_ALIGN, false>(g, b, fc, stream);
        if (g.solo) launch_inst<MODE, TW, MODE == MODE_SW_ALIGN, true>(g, b, fc, stream);
        return;
    }
    if (g.n >= 2) launch_inst<MODE, TW, false, false>(g, b, fc, stream);
    if (g.solo) launch_inst<MODE, TW, false, true>(g, b, fc, stream);
}

template <int MODE>
void launch_tw(const ChunkGeom &g, const ChunkBuffers &b, const FastConsts &fc, cudaStream_t stream) {
    if constexpr (MODE == MODE_SW_ALIGN) {  // more live state per cell: narrower strips keep it in registers
        if (g.policy == 1) {  // SSE/AVX pointers: 16-column strips only (fast_pick_tw)
            if (fc.gF == fc.gR) {
                if (g.n >= 2) launch_inst<MODE, 16, true, false, 1>(g, b, fc, stream);
                if (g.solo) launch_inst<MODE, 16, true, true, 1>(g, b, fc, stream);
            } else {
                if (g.n >= 2) launch_inst<MODE, 16, false, false, 1>(g, b, fc, stream);
                if (g.solo) launch_inst<MODE, 16, false, true, 1>(g, b, fc, stream);
            }
            return;
        }
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

// inband: the entry is 4 * s + 2 -- the value scale of the tagged NW align recurrence and its DIAG tag (va_nw.cu)
uint32_t table_word(int code, int match, int mismatch, int offset, bool inband = false) {
    uint32_t w = 0;
    for (int f = 0; f < 4; ++f) {
        int s = (code < 4 ? (code == f ? match : mismatch) : 0) - offset;
        if (inband) s = 4 * s + 2;
        w |= ((uint32_t)s & 0xFFu) << (8 * f);
    }
    return w;
}

bool fits8(int v) { return v >= -128 && v <= 127; }

}  // namespace

// The packed kernels are exact only while (a) every table entry fits a signed byte and (b) no
// cell can leave the int16 range; otherwise the call stays on the general 32-bit kernel.
bool fast_scoring_ok(int mode, int policy, const Scoring &sc, int read_length, int ref_length, bool intra) {
    const bool align = mode == MODE_NW_ALIGN || mode == MODE_SW_ALIGN;
    // SSE/AVX pointer rule: the second plane records LEFT >= UP (va_nw.cu, va_fast.cu); SW align under that rule walks
    // through zero cells and stores a third plane -- inter-task kernel only (the intra-task one has no room for it yet)
    if (align && policy != 0 && mode == MODE_SW_ALIGN && intra) return false;
    int mx = 1;
    for (int v : {sc.match, sc.mismatch, sc.gap_read, sc.gap_ref}) mx = max(mx, v < 0 ? -v : v);
    // the most one diagonal step can add: a "mismatch" score above the match score counts too
    const long long gain = max(max(sc.match, sc.mismatch), 0);
    const long long min_len = read_length < ref_length ? read_length : ref_length;
    if (intra) {
        // va_intra.cu: every mode in the Smith-Waterman form (tables s or s - gap_ref, values carried as "H + gap"), so
        // the range is the matrix's own.  Above: gain * min(len).  Below: SW cells are floored at 0; an NW score cell is
        // at least max(i*gap_ref, j*gap_read) (straight down / right from the zero borders), an NW align cell at least
        // i*gap_ref (straight down from row 0).
        if (sc.gap_read > 0 || sc.gap_ref > 0) return false;
        const int off = align ? sc.gap_ref : 0;
        if (!fits8(sc.match - off) || !fits8(sc.mismatch - off) || !fits8(-off)) return false;
        long long low = 0;
        if (mode == MODE_NW_SCORE) low = (long long)mx * (min_len + 2);
        if (mode == MODE_NW_ALIGN) low = -(long long)sc.gap_ref * (read_length + 2);
        return gain * min_len + 2 * mx <= 32000 && low + 2 * mx <= 32000 && mx <= 8000;
    }
    if (mode == MODE_NW_SCORE || mode == MODE_NW_ALIGN) {
        // shifted recurrence (va_nw.cu): gap scores <= 0, table entries s - gap_ref - gap_read, and
        // 0 <= V <= gain*min(rows,cols) + |gap_ref|*rows + |gap_read|*cols
        if (sc.gap_read > 0 || sc.gap_ref > 0) return false;
        const int off = sc.gap_ref + sc.gap_read;
        if (!fits8(sc.match - off) || !fits8(sc.mismatch - off) || !fits8(-off)) return false;
        const long long top = gain * min_len - (long long)sc.gap_ref * (read_length + 2) - (long long)sc.gap_read * (ref_length + 2);
        return top + 256 <= 32000;
    }
    if (sc.gap_read > 0 || sc.gap_ref > 0) return false;  // the columns past n of a partial strip rely on it
    const int off = align ? sc.gap_ref : 0;
    if (!fits8(sc.match - off) || !fits8(sc.mismatch - off) || !fits8(-off)) return false;
    // every cell is floored at 0: values stay within [-mx, gain * min(rows, cols)]
    const long long top = gain * min_len;
    // SW align packs (value + bias) * 32 + column into a 16-bit lane: value + 128 + gap must stay below 1024
    if (mode == MODE_SW_ALIGN) return top + 2 * 128 < 1024 && mx <= 100;
    return top + mx <= 32000 && mx <= 8000;
}

// The tagged form of the packed NW align kernel (va_nw.cu) carries 4V + tag in the 16-bit lanes: a quarter of the value
// range, and table entries 4s' + 2 that still have to fit a signed byte.  Calls outside that run the plane form.
bool fast_inband_ok(int mode, int policy, const Scoring &sc, int read_length, int ref_length) {
    if (mode != MODE_NW_ALIGN || !fast_scoring_ok(mode, policy, sc, read_length, ref_length)) return false;
    static const bool off_by_env = [] { const char *v = getenv("VERSALIGN_CUDA_NO_INBAND"); return v && atoi(v) != 0; }();
    if (off_by_env) return false;
    const int off = sc.gap_ref + sc.gap_read;
    for (int s : {sc.match - off, sc.mismatch - off, -off})
        if (!fits8(4 * s + 2)) return false;
    const long long gain = max(max(sc.match, sc.mismatch), 0);
    const long long min_len = read_length < ref_length ? read_length : ref_length;
    const long long top = gain * min_len - (long long)sc.gap_ref * (read_length + 2) - (long long)sc.gap_read * (ref_length + 2);
    return 4 * (top + 64) <= 32000;
}

int fast_pick_tw(int mode, int ref_length, int policy) {
    if (mode == MODE_SW_ALIGN && policy == 1) return 16;  // three planes: one group of 16 columns
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

FastConsts make_fast_consts(int mode, const Scoring &sc, bool inband) {
    FastConsts fc{};
    const bool align = mode == MODE_NW_ALIGN || mode == MODE_SW_ALIGN;
    const bool nw = mode == MODE_NW_SCORE || mode == MODE_NW_ALIGN;  // shifted recurrence, va_nw.cu
    const int off = nw ? sc.gap_ref + sc.gap_read : align ? sc.gap_ref : 0;
    inband = inband && mode == MODE_NW_ALIGN;
    for (int c = 0; c < 8; ++c) fc.tab[c] = table_word(c, sc.match, sc.mismatch, off, inband);
    // rows in front of a late-starting lane: s' = 0 hands matrix row 0 down (va_nw.cu); tagged form: 4 * 0 + 2
    if (nw) fc.tab[CODE_PRE] = inband ? 0x02020202u : 0u;
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
        fc.swa_k32 = 32;
        fc.swa_zero = pk(B);
    }
    return fc;
}

int launch_fill_fast(const ChunkGeom &g, const ChunkBuffers &b, int mode, const Scoring &sc, cudaStream_t stream) {
    if (g.fast_tw == 0 || (g.n < 2 && !g.solo)) return 0;
    const FastConsts fc = make_fast_consts(mode, sc, g.inband != 0);
    switch (mode) {
        case MODE_SW_SCORE: launch_tw<MODE_SW_SCORE>(g, b, fc, stream); break;
        case MODE_NW_SCORE:
        case MODE_NW_ALIGN: return launch_fill_nw(g, b, mode, fc, stream);
        case MODE_SW_ALIGN: launch_tw<MODE_SW_ALIGN>(g, b, fc, stream); break;
        default: return 0;
    }
    // kernels launched: the duo kernel (batches of two pairs and more) and the solo kernel
    return (g.n >= 2 ? 1 : 0) + (g.solo ? 1 : 0);
}

}  // namespace va
