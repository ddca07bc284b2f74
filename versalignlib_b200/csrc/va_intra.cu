// va_intra.cu -- intra-task kernels for long pairs: a CTA computes ONE pair-of-pairs (two pairs in the two
// s16 lanes of every register); each of its warps walks a 512-column pass of the matrix as a skewed
// wavefront and the warps pipeline consecutive passes.  The reference has nothing like it -- all of its
// kernels walk one matrix row-major with a loop-carried dependency through memory -- so parity
// is defined by the recurrences alone (score: DefaultKernel.cpp:83-202; fill with pointers:
// DefaultKernel.cpp:204-389, SSEKernel.cpp:646-659 for the SSE/AVX tie order): the cell arithmetic is the
// packed kernels' (va_fast.cu), only the order of evaluation changes.
//
//   * lane l owns 16 consecutive ref columns of the current pass (512 columns per pass);
//   * at step t lane l computes row t-l of its columns, so its left neighbour's right edge
//     (computed one step earlier) arrives by __shfl_up_sync together with that row's two score tables; lane 0
//     takes its inputs from a per-32-steps batch that all lanes load coalesced (row index -> tables, previous
//     pass's boundary column);
//   * lane 31's right edge is collected over 32 steps and stored coalesced as the next pass's
//     boundary column;
//   * the passes of one duo are spread over the warps of the CTA and run as a pipeline (see the kernel);
//   * align modes: the two comparisons of the pointer rule leave the max instructions as predicates and are
//     banked into bit planes by predicated FADDs exactly like the inter-task kernels; a row's planes of a
//     lane's 16 columns are one 8-byte word, stored at [duo][16-column strip][row] -- consecutive rows are
//     consecutive words, which is what the warp-cooperative traceback of long pairs reads 32 rows at a time
//     (va_traceback.cu);
//   * all four modes share the Smith-Waterman form of the recurrence (values carried as "H + gap"); the
//     Needleman-Wunsch modes drop the zero floor and differ at the borders.  (The shifted NW recurrence of
//     va_nw.cu would leave the 16-bit range on long pairs: its values grow with rows + columns.)
//   * SW align's best cell (first strictly greater in row-major order, DefaultKernel.cpp:252-256) cannot use
//     the inter-task kernel's 16-bit (value, column) key -- long pairs score far above 1023 -- so each lane
//     tracks the greatest row maximum of its strips and, whenever one of its halves sets a new one, parks the
//     row's 16 registers in shared memory; the column is read from that snapshot at the end.
// Used when a batch has too few pairs to fill the GPU with one thread per pair-of-pairs and
// the pairs are long (intra_preferred decides).
#include <algorithm>
#include <type_traits>

#include "va_fast.cuh"

namespace va {

namespace {

constexpr uint32_t NEG2 = 0x80008000u;
__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { return __viaddmax_s16x2(a, b, NEG2); }
__device__ __forceinline__ uint32_t pk(int v) { return ((uint32_t)v & 0xFFFFu) * 0x00010001u; }
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(int16_t)(v & 0xFFFFu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)(int16_t)(v >> 16); }

constexpr int MAX_PASSES = 64;  // 32 000 columns / (32 lanes x 16 columns)
constexpr int TW = INTRA_TW;
constexpr int MAX_WARPS = 16;
// register cap: 128 = 16 warps per SM (four per scheduler); the launcher sizes CTAs from it
#ifndef VA_INTRA_MAXREG
#define VA_INTRA_MAXREG 128
#endif

// One CTA = one duo; its W warps take the column passes round-robin (warp w: passes w, w+W, ...) and run
// them as a pipeline: pass p+1 follows pass p a few dozen rows behind, reading the boundary column pass p
// leaves in global memory (32 rows per coalesced store) as soon as a shared-memory progress counter says
// those rows are there.  A batch of a few long pairs then fills the GPU with W times as many warps.
//
// MODE: one of the four functions.  POLICY (NW align): which comparison the second plane records (0: UP >= LEFT,
// 1: LEFT >= UP).  SYM (align): gap_read == gap_ref, so "H + gR" and "H + gF" are one register.
template <int MODE, int POLICY, bool SYM>
__global__ void __maxnreg__(VA_INTRA_MAXREG) fill_intra_kernel(ChunkGeom g, ChunkBuffers b, FastConsts fc, int duo_first) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool ALIGN = MODE == MODE_SW_ALIGN || MODE == MODE_NW_ALIGN;
    constexpr bool SW = MODE == MODE_SW_SCORE || MODE == MODE_SW_ALIGN;
    constexpr bool SWA = MODE == MODE_SW_ALIGN, NWA = MODE == MODE_NW_ALIGN, SWS = MODE == MODE_SW_SCORE, NWS = MODE == MODE_NW_SCORE;
    __shared__ uint2 s_T2[64];            // [7*code_a + code_b] -> the two lanes' 4-entry score tables (49 used)
    __shared__ int prog[MAX_PASSES];      // rows of pass p whose right edge is in global memory
    __shared__ uint32_t warp_best[MAX_WARPS];
    __shared__ uint32_t s_edge[MAX_WARPS][64];  // lane 31's right edges of the last (up to) 64 rows, per warp
    __shared__ int warp_cell[MAX_WARPS][2][3];  // SW align: per warp and half: value, row, column of its best cell
    extern __shared__ uint32_t s_snap[];  // SW align: [warp][slot 0/1][register][lane] row snapshots
    if (threadIdx.x < 64) s_T2[threadIdx.x] = threadIdx.x < 49 ? make_uint2(fc.tab[threadIdx.x / 7], fc.tab[threadIdx.x % 7]) : make_uint2(0u, 0u);
    for (int t = threadIdx.x; t < MAX_PASSES; t += blockDim.x) prog[t] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int duo = duo_first + (int)blockIdx.x;
    const int slot_a = 2 * duo, slot_b = slot_a + 1;
    if (slot_b >= g.n) return;  // CTA-uniform
    const PairMeta ma = b.meta[slot_a], mb = b.meta[slot_b];
    if (!duo_is_fast(g, MODE, slot_a, ma, mb)) return;
    const int m = max((int)ma.rows, (int)mb.rows), n = ma.cols;
    const int gF = fc.gF, gR = fc.gR;
    const uint32_t gF2 = fc.gF2, gR2 = fc.gR2, dFR2 = fc.dFR2;  // dFR2: boundary ("H + gR") -> diagonal form
    const uint8_t *ridx = reinterpret_cast<const uint8_t *>(b.row_idx) + (size_t)duo * 16;
    const uint8_t *cref = reinterpret_cast<const uint8_t *>(b.code_refs);
    const uint32_t chunk_stride = (uint32_t)g.slots * 16u;
    const uint32_t ridx_stride = (uint32_t)g.duos * 16u;
    uint32_t *bnd = b.fboundary + (size_t)duo * g.rows_alloc;  // this duo's boundary column (rewritten pass after pass)
    volatile int *vprog = prog;
    const int ns = g.ref_chunks;
    const size_t rows2 = (size_t)intra_dir_rows(g);
    uint2 *dirs2 = reinterpret_cast<uint2 *>(b.fdirs) + (size_t)duo * ns * rows2;

    uint32_t best = 0;  // score modes: running maximum (SW) / max(0, last column, last row) (NW)
    // SW align: greatest row maximum so far per half, in the registers' form (value + gF); row, strip and
    // snapshot slot it came from
    uint32_t gbest2 = gF2, floor2 = gF2;
    int gi_a = 0, gi_b = 0, gs_a = -1, gs_b = -1, src_a = 0, src_b = 0;
    uint32_t *snap = s_snap + (size_t)warp * 2 * TW * 32 + lane;  // [slot][register] at stride 32

    const int pass_cols = 32 * TW;
    const int npasses = (n + pass_cols - 1) / pass_cols;
    const int steps = (m + 1) / 2 + 31;  // two rows per step
    const int last_lane = ((n - 1) >> 4) & 31;
    const int ra_last = (int)ma.rows - 1, rb_last = (int)mb.rows - 1;
    for (int pass = warp; pass < npasses; pass += W) {
        const int c_base = pass * pass_cols;
        const bool first_pass = pass == 0, last_pass = pass == npasses - 1, last_pass_rt = last_pass;
        const int c0 = c_base + lane * TW;
        const int kv = min(TW, max(0, n - c0));  // my valid columns in this pass
        const int kv_lane = kv;
        const int strip = c0 >> 4;
        uint32_t sel[TW], H[TW], H2[TW];
#pragma unroll
        for (int k = 0; k < TW; ++k) {
            const int col = min(c0 + k, n - 1);
            const size_t off = (size_t)(col >> 4) * chunk_stride + (col & 15);
            const uint32_t fa = cref[off + (size_t)slot_a * 16], fb = cref[off + (size_t)slot_b * 16];
            // columns past n: a selector that yields s <= 0 for both lanes (the sign byte of a table entry)
            sel[k] = k < kv ? (fa | ((fa | 8u) << 4) | ((fb | 4u) << 8) | ((fb | 12u) << 12)) : 0xCC88u;
            H[k] = ALIGN ? gF2 : 0u;  // matrix row 0 is 0 (align keeps H + gF)
        }
        uint32_t diag_next = ALIGN ? gF2 : 0u;            // H[row][c0] of the previous row, 0 for matrix row 0
        // what this lane used / produced for the two rows of its last step
        uint32_t cur_ta0 = 0, cur_tb0 = 0, edge0 = gR2, cur_ta1 = 0, cur_tb1 = 0, edge1 = gR2;
        uint2 *dp = dirs2 + (size_t)strip * rows2;
        uint2 w_even = make_uint2(0u, 0u);  // direction word of the even row of the step

        // One matrix row of this lane's 16 columns: Hi = the row above, H = this row (two register sets: a row's new
        // H[k] cannot overwrite the old one while the next cell still needs it as its diagonal).
        // full_tag: compile-time "every lane of this pass has all 16 columns and the pass is not the last" -- true for the
        // steady-state batches of all passes but the last, where it removes the partial-strip and last-pass tests.
        auto do_row = [&](auto full_tag, const int row, const uint32_t ta, const uint32_t tb, uint32_t left, const uint32_t(&Hi)[TW],
                          uint32_t(&H)[TW], uint32_t &edge_out) {
            constexpr bool FULL = decltype(full_tag)::value;
            const int kv = FULL ? TW : kv_lane;
            const bool last_pass = FULL ? false : last_pass_rt;
                    uint32_t diag = diag_next;
                    diag_next = add2(left, dFR2);  // "H + gR" -> the diagonal's form (H in the score modes, H + gF in the align modes)
                    if (ALIGN) {
                        float p1l = 8388608.0f, p1h = 8388608.0f, p2l = 8388608.0f, p2h = 8388608.0f;
#pragma unroll
                        for (int k = 0; k < TW; ++k) {
                            const uint32_t sub = prmt(ta, tb, sel[k]);  // table holds s - gF
                            const uint32_t up = Hi[k];
                            bool dl, dh, ul, uh;
                            // up+gF vs left+gR; policy 0: UP before LEFT, policy 1: LEFT before UP
                            const uint32_t tmax = POLICY == 0 ? __vibmax_s16x2(up, left, &uh, &ul) : __vibmax_s16x2(left, up, &uh, &ul);
                            const uint32_t d = add2(diag, sub);
                            const uint32_t h = __vibmax_s16x2(d, tmax, &dh, &dl);  // diag+s >= max(up,left): DIAG first
                            const float bit = (float)(1u << k);
                            if (dl) p1l += bit;
                            if (dh) p1h += bit;
                            if (ul) p2l += bit;
                            if (uh) p2h += bit;
                            // SW: the pointer of a positive cell is this rule; a zero cell is START, which the traceback
                            // recognises by tracking the score (DefaultKernel.cpp:238-248)
                            left = SW ? __viaddmax_s16x2(h, gR2, gR2) : add2(h, gR2);                  // max(h, 0) + gR
                            H[k] = SYM ? left : (SW ? __viaddmax_s16x2(h, gF2, gF2) : add2(h, gF2));  // max(h, 0) + gF
                            diag = up;
                        }
                        {
                            // a step computes an even row and the odd row after it: the even row keeps its word, the odd row
                            // stores both as one aligned 16-byte word (8-byte stores of single rows cost the L2 twice the
                            // partial-sector writes: 2.2 -> 3.1 TCUPS; one 256-bit store per four rows measured no better).
                            // The last row of an odd-sized matrix goes out alone.
                            uint2 w;
                            w.x = __byte_perm(__float_as_uint(p1l), __float_as_uint(p1h), 0x5410);
                            w.y = __byte_perm(__float_as_uint(p2l), __float_as_uint(p2h), 0x5410);
                            if (kv > 0) {
                                if (row & 1) *reinterpret_cast<uint4 *>(dp + row - 1) = make_uint4(w_even.x, w_even.y, w.x, w.y);
                                else if (row == m - 1) dp[row] = w;
                            }
                            w_even = w;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < TW; ++k) {
                            const uint32_t sub = prmt(ta, tb, sel[k]);
                            const uint32_t up = Hi[k];
                            const uint32_t tt = __viaddmax_s16x2(up, gF2, left);
                            const uint32_t h = SW ? __viaddmax_s16x2_relu(diag, sub, tt) : __viaddmax_s16x2(diag, sub, tt);
                            left = add2(h, gR2);
                            H[k] = h;
                            diag = up;
                        }
                    }
                    edge_out = left;
                    if (SWS) {
                        if (kv == TW) {
#pragma unroll
                            for (int k = 1; k < TW; k += 2) best = __vimax3_s16x2(best, H[k], H[k - 1]);
                        } else {
#pragma unroll
                            for (int k = 0; k < TW; ++k)
                                if (k < kv) best = __vmaxs2(best, H[k]);  // columns past n stay out of the maximum
                        }
                    }
                    if (SWA && kv > 0) {
                        // the row's maximum over my (valid) columns, both halves at once
                        uint32_t rm = NEG2;
                        if (kv == TW) {
#pragma unroll
                            for (int k = 1; k < TW; k += 2) rm = __vimax3_s16x2(rm, H[k], H[k - 1]);
                        } else {
#pragma unroll
                            for (int k = 0; k < TW; ++k)
                                if (k < kv) rm = __vmaxs2(rm, H[k]);
                        }
                        // a new best: strictly greater than the best so far -- or equal to it in an EARLIER row, which only a
                        // later pass can meet (rows above the row of the best so far)
                        bool keep_b, keep_a, fl_b, fl_a;
                        (void)__vibmax_s16x2(gbest2, rm, &keep_b, &keep_a);  // best so far >= row maximum?
                        // ... and at least the warp's best as of the last batch (`floor2`): on similar sequences every lane
                        // inside the band around the main diagonal sets a new LOCAL best in almost every row, but only
                        // values that reach the warp-wide best can be the pair's best cell
                        (void)__vibmax_s16x2(rm, floor2, &fl_b, &fl_a);
                        bool cand = (!keep_a && fl_a) || (!keep_b && fl_b);
                        if (row < max(gi_a, gi_b)) {
                            bool ge_b, ge_a;
                            (void)__vibmax_s16x2(rm, gbest2, &ge_b, &ge_a);
                            cand = cand || (ge_a && fl_a) || (ge_b && fl_b);
                        }
                        if (cand) {
                            const int ra = lo16(rm), rb = hi16(rm), ba = lo16(gbest2), bb = hi16(gbest2);
                            // strictly greater, or equal in an earlier row (a later pass revisits earlier rows)
                            const bool new_a = (ra > ba || (ra == ba && gs_a >= 0 && row < gi_a)) && fl_a;
                            const bool new_b = (rb > bb || (rb == bb && gs_b >= 0 && row < gi_b)) && fl_b;
                            if (new_a || new_b) {
                                // park the row in the slot the other half's snapshot does not live in
                                const int slot = (new_a && new_b) ? 0 : (new_a ? (src_b ^ 1) : (src_a ^ 1));
                                uint32_t *sp = snap + slot * (TW * 32);
#pragma unroll
                                for (int k = 0; k < TW; ++k) sp[k * 32] = H[k];
                                if (new_a) {
                                    gbest2 = (gbest2 & 0xFFFF0000u) | (rm & 0x0000FFFFu);
                                    gi_a = row;
                                    gs_a = strip;
                                    src_a = slot;
                                }
                                if (new_b) {
                                    gbest2 = (gbest2 & 0x0000FFFFu) | (rm & 0xFFFF0000u);
                                    gi_b = row;
                                    gs_b = strip;
                                    src_b = slot;
                                }
                            }
                        }
                    }
                    if (!SW && last_pass && lane == last_lane) {
                        // the last true column of this row: register kv-1 of this lane
                        uint32_t lc = H[TW - 1];
#pragma unroll
                        for (int k = 0; k < TW - 1; ++k)
                            if (kv == k + 1) lc = H[k];
                        if (NWS) best = __vmaxs2(best, lc);                          // SSEKernel.cpp:1285-1291
                        else bnd[row] = add2(lc, pk(-gF));                           // true H: the traceback's pad-column rule reads it
                    }
                    if (NWS && row == m - 1 && kv > 0) {  // whole last row (SSEKernel.cpp:1302-1310)
#pragma unroll
                        for (int k = 0; k < TW; ++k)
                            if (k < kv) best = __vmaxs2(best, H[k]);
                    }
                    if (NWA && kv > 0 && (row == ra_last || row == rb_last)) {
                        // the row the end-cell rule scans (DefaultKernel.cpp:352-355,381-387): per strip the best of its
                        // columns as one key per lane, (H << 16) | (0xFFFF - column); the
                        // traceback kernel reduces the strip keys of a pair (va_nw.cu's format, un-shifted values)
                        int key_a = (int)0x80000000, key_b = (int)0x80000000;
                        const uint32_t un = pk(-gF);  // registers hold H + gF; the keys carry the true H (long pairs: H - gap_ref*rows would leave 16 bits)
#pragma unroll
                        for (int k = 0; k < TW; ++k) {
                            if (k < kv) {
                                const uint32_t cand = add2(H[k], un);
                                const uint32_t low = 0xFFFFu - (uint32_t)(c0 + k);
                                key_a = max(key_a, (int)((cand << 16) | low));
                                key_b = max(key_b, (int)((cand & 0xFFFF0000u) | low));
                            }
                        }
                        uint32_t *hk = b.hrow + ((size_t)strip * g.duos + duo) * 2;
                        if (row == ra_last) hk[0] = (uint32_t)key_a;
                        if (row == rb_last) hk[1] = (uint32_t)key_b;
                    }
        };

        // A step = TWO rows per lane: at step t lane l computes rows 2(t-l) and 2(t-l)+1, so both rows' left edges and
        // tables come from the left neighbour's previous step.  The two rows are independent enough for the scheduler to
        // interleave them (row r+1 may start as soon as row r has its first cells), which halves the dependent chain per
        // cell that a warp exposes between two shuffles.
        for (int t0 = 0; t0 < steps; t0 += 32) {
            // batch inputs of lane 0 for steps t0..t0+31: rows r0 = 2 * (t0 + lane) and r0 + 1
            const int r0 = 2 * (t0 + lane);
            // matrix column 0 as "H + gR": 0 everywhere but NW align, where H(I,0) = I*gap_ref (DefaultKernel.cpp:304)
            uint32_t bat_a0 = 0, bat_b0 = 0, bat_l0 = NWA ? pk((r0 + 1) * gF + gR) : gR2;
            uint32_t bat_a1 = 0, bat_b1 = 0, bat_l1 = NWA ? pk((r0 + 2) * gF + gR) : gR2;
            if (!first_pass && 2 * t0 < m) {
                // the previous pass (another warp of this CTA when W > 1) must have left these rows
                const int need = min(2 * t0 + 64, m);
                while (vprog[pass - 1] < need) {
                }
                __threadfence_block();
            }
            if (r0 < m) {
                const uint2 tt2 = s_T2[ridx[(size_t)(r0 >> 4) * ridx_stride + (r0 & 15)]];
                bat_a0 = tt2.x;
                bat_b0 = tt2.y;
                if (!first_pass) bat_l0 = __ldcg(bnd + r0);  // written by another warp: read it where it was written (L2)
            }
            if (r0 + 1 < m) {
                const uint2 tt2 = s_T2[ridx[(size_t)((r0 + 1) >> 4) * ridx_stride + ((r0 + 1) & 15)]];
                bat_a1 = tt2.x;
                bat_b1 = tt2.y;
                if (!first_pass) bat_l1 = __ldcg(bnd + r0 + 1);
            }
            __syncwarp();  // every lane has read its rows before this warp overwrites the column below
            if (SWA) {  // the warp's best so far, per half (the snapshot filter above)
                floor2 = gbest2;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) floor2 = __vmaxs2(floor2, __shfl_xor_sync(FULL, floor2, o));
            }
            const int s_end = min(32, steps - t0);
            auto do_step = [&](auto all_rows_valid, const int s) {
                const int t = t0 + s;
                // lane 0 reads the batch, every other lane takes what its left neighbour used last step
                const uint32_t a0 = __shfl_sync(FULL, bat_a0, s), b0 = __shfl_sync(FULL, bat_b0, s), l0 = __shfl_sync(FULL, bat_l0, s);
                const uint32_t a1 = __shfl_sync(FULL, bat_a1, s), b1 = __shfl_sync(FULL, bat_b1, s), l1 = __shfl_sync(FULL, bat_l1, s);
                const uint32_t pa0 = __shfl_up_sync(FULL, cur_ta0, 1), pb0 = __shfl_up_sync(FULL, cur_tb0, 1), pl0 = __shfl_up_sync(FULL, edge0, 1);
                const uint32_t pa1 = __shfl_up_sync(FULL, cur_ta1, 1), pb1 = __shfl_up_sync(FULL, cur_tb1, 1), pl1 = __shfl_up_sync(FULL, edge1, 1);
                const int row = 2 * (t - lane);
                if (decltype(all_rows_valid)::value || (row >= 0 && row < m)) {
                    cur_ta0 = lane == 0 ? a0 : pa0;
                    cur_tb0 = lane == 0 ? b0 : pb0;
                    do_row(all_rows_valid, row, cur_ta0, cur_tb0, lane == 0 ? l0 : pl0, H, H2, edge0);
                } else {
#pragma unroll
                    for (int k = 0; k < TW; ++k) H2[k] = H[k];
                }
                if (decltype(all_rows_valid)::value || (row >= 0 && row + 1 < m)) {
                    cur_ta1 = lane == 0 ? a1 : pa1;
                    cur_tb1 = lane == 0 ? b1 : pb1;
                    do_row(all_rows_valid, row + 1, cur_ta1, cur_tb1, lane == 0 ? l1 : pl1, H2, H, edge1);
                } else {
#pragma unroll
                    for (int k = 0; k < TW; ++k) H[k] = H2[k];
                }
                // lane 31 finished rows 2(t-31), 2(t-31)+1: it parks their right edges in shared memory; every 64 rows the
                // warp stores them coalesced
                if (!last_pass) {
                    const int orow = 2 * (t - 31);
                    if (orow >= 0) {
                        if (lane == 31) {
                            s_edge[warp][orow & 63] = edge0;
                            s_edge[warp][(orow & 63) + 1] = edge1;
                        }
                        if ((orow & 63) == 62 || orow + 2 >= m) {
                            __syncwarp();
                            const int row0 = orow & ~63;
                            if (row0 + lane < m && row0 + lane <= orow + 1) bnd[row0 + lane] = s_edge[warp][lane];
                            if (row0 + 32 + lane < m && row0 + 32 + lane <= orow + 1) bnd[row0 + 32 + lane] = s_edge[warp][32 + lane];
                            __threadfence_block();
                            __syncwarp();
                            if (lane == 0) vprog[pass] = min(orow + 2, m);  // rows [0, orow + 1] of this pass are out
                        }
                    }
                }
            };
            // Batches in which every lane has both rows (all but the first and the last one or two of a pass) of a pass in
            // which every lane has all its columns (all but the last) run a copy of the step without the row / column tests: with them, the skipped path pins H to its input registers and the
            // computed path pays a register move per cell to get there.
            if (t0 >= 32 && 2 * (t0 + 32) <= m && !last_pass) {
                for (int s = 0; s < 32; ++s) do_step(std::true_type{}, s);
            } else {
                for (int s = 0; s < s_end; ++s) do_step(std::false_type{}, s);
            }
        }
    }
    const unsigned long long cells = ((unsigned long long)ma.rows + (unsigned long long)mb.rows) * (unsigned long long)n;
    if (!ALIGN) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = __vmaxs2(best, __shfl_xor_sync(FULL, best, o));
        if (lane == 0) warp_best[warp] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < W; ++w) best = __vmaxs2(best, warp_best[w]);
            b.scores[b.pair_of[slot_a]] = (int16_t)(best & 0xFFFF);
            b.scores[b.pair_of[slot_b]] = (int16_t)(best >> 16);
            atomicAdd(b.cell_count, cells);
        }
    } else if (SWA) {
        // each lane: column of its best cell = first register of the snapshot that holds the value
        int va = lo16(gbest2) - gF, vb = hi16(gbest2) - gF, ja = 0, jb = 0;
        if (gs_a >= 0) {
            const uint32_t *sp = snap + src_a * (TW * 32);
            int k = 0;
            while (k < TW - 1 && lo16(sp[k * 32]) != lo16(gbest2)) ++k;
            ja = gs_a * 16 + k;
        } else {
            va = 0;
        }
        if (gs_b >= 0) {
            const uint32_t *sp = snap + src_b * (TW * 32);
            int k = 0;
            while (k < TW - 1 && hi16(sp[k * 32]) != hi16(gbest2)) ++k;
            jb = gs_b * 16 + k;
        } else {
            vb = 0;
        }
        // greatest value, then smallest row, then smallest column = first strictly greater cell in row-major order
        auto better = [](int v, int i, int j, int ov, int oi, int oj) { return ov > v || (ov == v && (oi < i || (oi == i && oj < j))); };
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int ova = __shfl_xor_sync(FULL, va, o), oia = __shfl_xor_sync(FULL, gi_a, o), oja = __shfl_xor_sync(FULL, ja, o);
            const int ovb = __shfl_xor_sync(FULL, vb, o), oib = __shfl_xor_sync(FULL, gi_b, o), ojb = __shfl_xor_sync(FULL, jb, o);
            if (better(va, gi_a, ja, ova, oia, oja)) { va = ova; gi_a = oia; ja = oja; }
            if (better(vb, gi_b, jb, ovb, oib, ojb)) { vb = ovb; gi_b = oib; jb = ojb; }
        }
        if (lane == 0) {
            warp_cell[warp][0][0] = va; warp_cell[warp][0][1] = gi_a; warp_cell[warp][0][2] = ja;
            warp_cell[warp][1][0] = vb; warp_cell[warp][1][1] = gi_b; warp_cell[warp][1][2] = jb;
        }
        __syncthreads();
        if (threadIdx.x < 2) {
            const int hf = threadIdx.x;
            int v = warp_cell[0][hf][0], i = warp_cell[0][hf][1], j = warp_cell[0][hf][2];
            for (int w = 1; w < W; ++w)
                if (better(v, i, j, warp_cell[w][hf][0], warp_cell[w][hf][1], warp_cell[w][hf][2])) {
                    v = warp_cell[w][hf][0]; i = warp_cell[w][hf][1]; j = warp_cell[w][hf][2];
                }
            if (v <= 0) v = i = j = 0;  // score 0: the traceback starts (and stops) at cell (0,0) (DefaultKernel.cpp:207-208)
            const int p = b.pair_of[hf ? slot_b : slot_a];
            b.scores[p] = (int16_t)v;
            b.end_cell[2 * p] = (int16_t)i;
            b.end_cell[2 * p + 1] = (int16_t)j;
            if (hf == 0) atomicAdd(b.cell_count, cells);
        }
    } else {
        if (threadIdx.x == 0) atomicAdd(b.cell_count, cells);
    }
}

template <int MODE, int POLICY, bool SYM>
void launch_intra_inst(const ChunkGeom &g, const ChunkBuffers &b, const FastConsts &fc, int duo_first, int duos, int warps, cudaStream_t stream) {
    const size_t smem = MODE == MODE_SW_ALIGN ? (size_t)warps * 2 * TW * 32 * sizeof(uint32_t) : 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(fill_intra_kernel<MODE, POLICY, SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fill_intra_kernel<MODE, POLICY, SYM><<<duos, 32 * warps, smem, stream>>>(g, b, fc, duo_first);
}

}  // namespace

// The intra-task kernels pay when one-thread-per-duo cannot fill the machine: few, long pairs.
bool intra_preferred(int mode, int n_pairs, int read_length, int ref_length, int sm_count) {
    (void)mode;
    const long long duos = n_pairs / 2;
    return ref_length >= 1024 && read_length >= 256 && duos < (long long)sm_count * 512;
}

// warps per duo (= per CTA) for a launch of `duos` CTAs: as many as there are column passes to pipeline; at least 4 when the
// passes allow it (8 measured no better when duos are plentiful), more when the launch has too few duos to fill the device's
// warp slots (16 warps per SM at this register budget)
static int intra_warps(int duos, int passes) {
    const int want = (148 * 16 + duos - 1) / std::max(duos, 1);
    int warps = std::max(4, std::min(want, MAX_WARPS));
    warps = std::min(warps, passes);
    if (warps >= 3 && warps != 4 && warps != 8 && warps != 16) warps = warps > 8 ? 8 : 4;  // 1, 2, 4, 8 or 16 warps per CTA
    return std::max(warps, 1);
}

int launch_fill_intra(const ChunkGeom &g, const ChunkBuffers &b, int mode, const FastConsts &fc, cudaStream_t stream) {
    if (g.n < 2) return 0;
    const int duos = g.n / 2;
    const int passes = (g.ref_length + 32 * TW - 1) / (32 * TW);
    const bool sym = fc.gF == fc.gR;
    auto launch = [&](int first, int count, int warps) {
        switch (mode) {
            case MODE_SW_SCORE: launch_intra_inst<MODE_SW_SCORE, 0, false>(g, b, fc, first, count, warps, stream); break;
            case MODE_NW_SCORE: launch_intra_inst<MODE_NW_SCORE, 0, false>(g, b, fc, first, count, warps, stream); break;
            case MODE_SW_ALIGN:
                if (sym) launch_intra_inst<MODE_SW_ALIGN, 0, true>(g, b, fc, first, count, warps, stream);
                else launch_intra_inst<MODE_SW_ALIGN, 0, false>(g, b, fc, first, count, warps, stream);
                break;
            default:
                if (g.policy == 1) {
                    if (sym) launch_intra_inst<MODE_NW_ALIGN, 1, true>(g, b, fc, first, count, warps, stream);
                    else launch_intra_inst<MODE_NW_ALIGN, 1, false>(g, b, fc, first, count, warps, stream);
                } else {
                    if (sym) launch_intra_inst<MODE_NW_ALIGN, 0, true>(g, b, fc, first, count, warps, stream);
                    else launch_intra_inst<MODE_NW_ALIGN, 0, false>(g, b, fc, first, count, warps, stream);
                }
                break;
        }
    };
    // Whole waves of 4-warp CTAs first (the device holds 4 per SM); what is left over gets its own launch with as many
    // warps per duo as fill the device -- as the tail of one launch those CTAs would run 4 warps each on a nearly empty
    // device for a whole wave's time (C4 at 8 GPUs: 625 duos per rank = 592 + 33).
    const int wave = 148 * 4;
    const int full = passes >= 4 && duos > wave ? duos / wave * wave : 0;
    int launches = 0;
    if (full) {
        launch(0, full, 4);
        ++launches;
    }
    if (duos > full) {
        launch(full, duos - full, intra_warps(duos - full, passes));
        ++launches;
    }
    return launches;
}

}  // namespace va
