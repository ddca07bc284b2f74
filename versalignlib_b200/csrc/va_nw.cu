// va_nw.cu -- the packed inter-task fill kernels of the library's "Needleman-Wunsch" modes
// (score: DefaultKernel.cpp:140-202, SSEKernel.cpp:1152-1315; fill with pointers:
// DefaultKernel.cpp:282-389), two pairs per thread in the two s16 lanes like va_fast.cu.
//
// What is different from the SW kernels: nothing in these modes clamps a cell at 0, so the whole
// matrix can be computed in SHIFTED coordinates
//        V(I,J) = H(I,J) - gap_ref*I - gap_read*J          (I,J = matrix row / column, 0-based)
// in which both gap moves cost nothing:
//        V(I,J) = max( V(I-1,J-1) + (s - gap_ref - gap_read),  V(I-1,J),  V(I,J-1) ).
// The three candidates of a cell carry the same shift, so every comparison the reference makes
// (and with it every traceback pointer) is unchanged, while the cell update shrinks to
//        score:  t = max(up, left);  V = max(diag + s', t)                 2 DPX instructions
//        align:  t = max(up, left) -> UP>=LEFT;  d = diag + s';  V = max(d, t) -> DIAG>=rest   3
// (+ 1 PRMT for s'), against 3 and 4 for the unshifted recurrence: "H + gap" never has to be
// formed, and the value a cell hands to its right neighbour and to the row below is one register.
// With both gap scores <= 0 (required here; otherwise the general kernel runs) V is non-negative and
// non-decreasing along rows and columns, which bounds the 16-bit range check.
// A duo whose reads differ in length is END-aligned in align mode: the shorter lane's first rows carry the
// row index CODE_PRE, whose table is all zero -- with s' = 0 a row reproduces the one above it (V(0,.) is
// non-decreasing), so matrix row 0 is handed down to where that lane really starts, both lanes finish on
// the last sweep row, and the end-cell rule reads both from the registers.
// TAGGED form of the align kernel (INBAND, used whenever 4x the value range still fits 16 bits -- every BASELINE
// config): the lanes carry 4V + tag, the table entries are 4s' + 2, and
//        t = max(first + 1, second)          first = UP (policy 0) / LEFT (policy 1): wins ties, tag 1; the other: tag 0
//        h = max(diag + (4s' + 2), t)        DIAG wins ties against both: tag 2
//        x = h & ~3                          the clean value every later cell reads
// so the two low bits of h ARE the traceback pointer and no comparison result has to leave an instruction as a
// predicate: 4 integer-pipe instructions per cell-pair like the plane form, but the four predicated FADDs that bank the
// planes become two IMADs (tag = h - x, word = word * 4 + tag) -- the plane form is bound by the issue port, this one by
// the integer pipe alone (tools/micro/cell_mix.cu).  A row's tags fill the same 16 bytes per pair-of-pairs
// as its two planes did: per group of 16 columns .x = columns 0..7, .y = columns 8..15, column c of a word at bits
// 2*(7 - c) (lane A) and 16 + 2*(7 - c) (lane B); same word addressing, the traceback decodes by ChunkGeom::inband.
// The matrix borders un-shift what they hand out: NW score's last row / last column maxima, the
// `hrow` row and the last-true-column values the traceback kernel reads (va_traceback.cu).
#include <algorithm>
#include <type_traits>

#include "va_fast.cuh"

namespace va {

namespace {

// Threads per block and register cap: 4 x 128 threads per SM leave 128 registers each, 5 x 96 leave 136.
// The row loop of the align kernels sits right at the limit, and ptxas' allocation there is not monotonic:
// under the wrong cap it parks the direction predicates in a register and re-materialises them with two
// LOP3 per cell (~540 -> ~900 instructions per row pair).  The pairs below are the ones whose loops come
// out clean (tools/check_sass.py asserts it at build time); score kernels are clean under both, 96 x 5
// measured faster.
template <bool ALIGN, int TW, bool SOLO>
struct Block {
    // align duo kernels: 4 x 128 threads under a cap of 128 registers; everything else 5 x 96 under 136
    static constexpr bool WIDE128 = ALIGN && !SOLO;
    static constexpr int NT = WIDE128 ? 128 : 96;
    static constexpr int MAXREG = WIDE128 ? 128 : 136;
};
constexpr uint32_t NEG2 = 0x80008000u;  // (-32768, -32768): identity of the packed max

__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { return __viaddmax_s16x2(a, b, NEG2); }
__device__ __forceinline__ uint32_t pk(int v) { return ((uint32_t)v & 0xFFFFu) * 0x00010001u; }

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

// SOLO: the instantiation for single slots whose duo is not fast (va_fast.cuh): same sweep, the owner's
// halves of the shared words stored with 16-bit stores.  Kept apart so the duo kernel's stores stay
// unconditional (its schedule sits right at the register budget).
// POLICY (align only): which comparison the second direction plane records -- 0: UP >= LEFT (Default / OpenCL,
// DefaultKernel.cpp:338-346: DIAG > UP > LEFT), 1: LEFT >= UP (SSE / AVX, SSEKernel.cpp:646-659: DIAG > LEFT > UP).
// The packed path only takes pairs whose swept rows and columns are all ACGT (va_fast.cuh, va_prep.cu), so the
// SSE/AVX rule's "DIAG only between two valid bases" never masks anything here and START cannot occur below row 0.
template <bool ALIGN, int TW, bool SOLO, int POLICY, bool INBAND>
__global__ void __maxnreg__((Block<ALIGN, TW, SOLO>::MAXREG)) fill_nw_kernel(ChunkGeom g, ChunkBuffers b, FastConsts fc) {
    static_assert(ALIGN || !INBAND, "the tagged form is an align form");
    constexpr int NG = (TW + 15) / 16;
    constexpr int VS = INBAND ? 4 : 1;  // value scale of the lanes
    constexpr int NT = Block<ALIGN, TW, SOLO>::NT;
    constexpr int MODE = ALIGN ? MODE_NW_ALIGN : MODE_NW_SCORE;

    __shared__ uint2 s_T2[256];              // [7*code_a + code_b] -> the two lanes' 4-entry score tables (49 used)
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
        const int ngF = -fc.gF, ngR = -fc.gR;  // both >= 0
        const uint32_t gF2 = fc.gF2, ngF2 = pk(ngF);
        const uint8_t *cref = reinterpret_cast<const uint8_t *>(b.code_refs);
        uint32_t *bnd = b.fboundary;
        uint4 *dirs = b.fdirs;

        uint32_t best = 0;  // score mode: max(0, last column, last row) of H
        const int nstrips = (n + TW - 1) / TW;
        for (int s = 0; s < nstrips; ++s) {
            const int c0 = s * TW;
            const bool first = s == 0, last = s == nstrips - 1;
            // A partial last strip keeps its kv true columns in the LAST kv registers; the `pad` registers in
            // front of them are pass-through columns: their selector yields s' <= 0 (the sign byte of a table
            // entry) and they start from V(0,c0), so with V non-decreasing down the boundary column each of
            // them reproduces its left neighbour, and the first true column sees exactly the boundary column
            // to its left and above-left.  The last true column is therefore always register TW-1, and the
            // sweep needs no per-column guards.  Their direction bits are never read (va_traceback.cu).
            const int kv = min(TW, n - c0);  // valid columns of this strip
            const int pad = TW - kv;
            uint32_t sel[TW], H[TW];
            // (the plane form of align keeps the byte-wise prologue: its row loop sits at a register-allocation cliff; the
            // score form measured 1 % slower with the lean one -- 96-thread blocks under another register cap)
            constexpr bool LEAN = INBAND;
            if (LEAN && pad == 0) {  // full strip, the common case
                fast_strip_selectors<TW>(g, b.code_refs, slot_a, c0, sel);
                const uint32_t hstep = pk(VS * ngR);
                uint32_t hv = pk(VS * ngR * c0);
#pragma unroll
                for (int k = 0; k < TW; ++k) {
                    hv += hstep;  // V(0,J) = -gap_read*J, both lanes (no carry between them: the range check)
                    H[k] = hv;
                }
            } else {  // partial strip (and the kernels that keep the byte-wise prologue)
#pragma unroll
                for (int k = 0; k < TW; ++k) {
                    const int col = max(c0 + k - pad, c0);
                    const size_t off = ((size_t)(col >> 4) * g.slots) * 16 + (col & 15);
                    const uint32_t fa = cref[off + (size_t)slot_a * 16], fb = cref[off + (size_t)slot_b * 16];
                    // nibbles: lane A low byte <- table a[fa], high byte <- its sign; lane B from table b (bytes 4..7)
                    sel[k] = k >= pad ? (fa | ((fa | 8u) << 4) | ((fb | 4u) << 8) | ((fb | 12u) << 12)) : 0xCC88u;
                    H[k] = pk(VS * ngR * (max(c0 + k - pad, c0 - 1) + 1));  // V(0,J) = -gap_read*J: matrix row 0 is 0
                }
            }
            uint32_t diag_next = pk(VS * ngR * c0);  // V(0, c0)
            // matrix column 0 in shifted form: H(I,0) = I*gap_ref -> 0 (align); H(I,0) = 0 -> -gap_ref*I (score)
            uint32_t col0 = ALIGN ? 0u : ngF2;
            // score mode, last strip: un-shift of the last column, gap_ref*I + gap_read*n
            uint32_t corr = pk(fc.gR * n + fc.gF);
            // One sweep over all read rows of this strip.
            //
            // Per-row inputs (the row's pair of substitution tables, the previous strip's right edge) never
            // make a row wait on HBM: they are staged 16 rows ahead with cp.async into this thread's own
            // shared-memory slots -- no registers held across the loop, nothing for the scheduler to sink
            // next to the consumer -- one commit group per 16-row chunk, double buffered.
            {
                uint32_t *bp = bnd + duo;
                uint4 *dp = dirs + fast_dir_index(g, s, 0, 0, duo);
                const int nchunks = (m + 15) >> 4;
                auto stage = [&](int c, int buf) {
                    cp_async16(&s_idx[buf][threadIdx.x], b.row_idx + (size_t)c * g.duos + duo);
                    if (!first) {
                        const uint32_t *src = bnd + (size_t)(c * 16) * g.duos + duo;
#pragma unroll
                        for (int r = 0; r < 16; ++r)
                            // (no row test -- the boundary block is allocated 16 rows past the last one, the rows staged in
                            // excess are never read back)
                            if (LEAN || c * 16 + r < m) cp_async4(&s_bnd[buf][r][threadIdx.x], src + (size_t)r * g.duos);
                    }
                    cp_async_commit();
                };
                // one matrix row of this strip; align returns the row's two direction planes per group
                auto do_row = [&](const uint2 tt, uint32_t left_in, uint2(&w)[NG]) {
                    const uint32_t ta = tt.x, tb = tt.y;
                    uint32_t left = first ? col0 : left_in;
                    if (!ALIGN && first) col0 = add2(col0, ngF2);
                    uint32_t diag = diag_next;
                    diag_next = left;  // V(I, c0) is the next row's diagonal
                    float p1l[NG], p1h[NG], p2l[NG], p2h[NG];
                    uint32_t tags[2 * NG];  // tagged form: 8 columns per word
#pragma unroll
                    for (int q = 0; q < NG; ++q) {
                        p1l[q] = p1h[q] = p2l[q] = p2h[q] = 8388608.0f;
                        tags[2 * q] = tags[2 * q + 1] = 0u;
                    }
#pragma unroll
                    for (int k = 0; k < TW; ++k) {
                        const uint32_t sub = prmt(ta, tb, sel[k]);  // s - gap_ref - gap_read
                        const uint32_t up = H[k];
                        uint32_t h;
                        if (INBAND) {
                            const uint32_t t = POLICY == 0 ? __viaddmax_s16x2(up, 0x00010001u, left) : __viaddmax_s16x2(left, 0x00010001u, up);
                            const uint32_t hh = __viaddmax_s16x2(diag, sub, t);
                            h = hh & 0xFFFCFFFCu;
                            // tag = hh - h, word = word * 4 + tag: both on the FMA pipe (as plain C they come back as
                            // integer-pipe adds and shifts)
                            uint32_t tag;
                            asm("mad.lo.u32 %0, %1, -1, %2;" : "=r"(tag) : "r"(h), "r"(hh));
                            asm("mad.lo.u32 %0, %0, 4, %1;" : "+r"(tags[k >> 3]) : "r"(tag));
                        } else if (ALIGN) {
                            bool dl, dh, ul, uh;
                            // policy 0: up >= left -> UP before LEFT; policy 1: left >= up -> LEFT before UP
                            const uint32_t t = POLICY == 0 ? __vibmax_s16x2(up, left, &uh, &ul) : __vibmax_s16x2(left, up, &uh, &ul);
                            const uint32_t d = add2(diag, sub);
                            h = __vibmax_s16x2(d, t, &dh, &dl);                     // diag >= max(up,left) : DIAG first
                            const float bit = (float)(1u << (k & 15));
                            if (dl) p1l[k >> 4] += bit;
                            if (dh) p1h[k >> 4] += bit;
                            if (ul) p2l[k >> 4] += bit;
                            if (uh) p2h[k >> 4] += bit;
                        } else {
                            h = __viaddmax_s16x2(diag, sub, __vmaxs2(up, left));
                        }
                        left = h;
                        H[k] = h;
                        diag = up;
                    }
                    // right edge of the strip for the next strip; align also keeps the LAST true column (the
                    // traceback kernel needs it to decide whether the padded arg-max lands in a pad column)
                    // (align: also of the last strip -- on a padded ref the traceback's pad-column rule needs the last true column)
                    if (ALIGN || !last) store_lanes<SOLO>(bp, left, fw);
                    bp += g.duos;
                    if (!ALIGN && last) {  // last column of this row, un-shifted (SSEKernel.cpp:1285-1291)
                        best = __viaddmax_s16x2(left, corr, best);
                        corr = add2(corr, gF2);
                    }
                    if (INBAND) {
                        // a strip's last word may hold fewer than 8 columns: left-align it, so that column c of every
                        // word sits at bits 2*(7 - c)
                        constexpr int LASTN = TW - 8 * ((TW - 1) / 8);
                        if (LASTN < 8) tags[(TW - 1) / 8] <<= 2 * (8 - LASTN);
#pragma unroll
                        for (int q = 0; q < NG; ++q) {
                            w[q].x = tags[2 * q];
                            w[q].y = tags[2 * q + 1];
                        }
                    } else if (ALIGN) {
#pragma unroll
                        for (int q = 0; q < NG; ++q) {
                            w[q].x = __byte_perm(__float_as_uint(p1l[q]), __float_as_uint(p1h[q]), 0x5410);
                            w[q].y = __byte_perm(__float_as_uint(p2l[q]), __float_as_uint(p2h[q]), 0x5410);
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
                    const int rend = min(16, m - c * 16);
                    // two rows per iteration: their direction words leave as one 16-byte store per group
                    int r = 0;
                    for (; r + 1 < rend; r += 2, dp += (size_t)NG * g.duos) {
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
                        do_row(t0, l0, w0);
                        if (ALIGN) {
#pragma unroll
                            for (int q = 0; q < NG; ++q) store_half<SOLO>(dp + (size_t)q * g.duos, 0, w0[q], fw);
                        }
                        do_row(t1, l1, w0);
                        if (ALIGN) {
#pragma unroll
                            for (int q = 0; q < NG; ++q) store_half<SOLO>(dp + (size_t)q * g.duos, 1, w0[q], fw);
                        }
                    }
                    if (r < rend) {  // odd row count (last chunk only): the last word holds one row
                        uint2 w0[NG];
                        do_row(nt0, nl0, w0);
                        if (ALIGN) {
#pragma unroll
                            for (int q = 0; q < NG; ++q) store_half<SOLO>(dp + (size_t)q * g.duos, 0, w0[q], fw);
                        }
                    }
                    buf = buf1;
                }
                cp_async_wait<0>();
            }
            // Pass-through columns hold the value of the column left of the strip, so they are handed out
            // like that column (no per-column guards: an idempotent store / a harmless repeat in the maximum).
            if (ALIGN) {
                // the row the end-cell rule scans (DefaultKernel.cpp:352-355,381-387), still shifted: the traceback
                // kernel finds its arg-max.  [column][duo]: coalesced across the warp.
                // per strip: best of its columns as one 32-bit key per lane, (value << 16) | (0xFFFF - column), value =
                // V + gap_read*J (= H - gap_ref*rows): the signed maximum is the larger value and, among equals, the
                // smaller column.  No predicates, nothing carried between strips: the traceback kernel reduces the
                // (at most a few dozen) strip keys of a pair.
                int key_a = (int)0x80000000, key_b = (int)0x80000000;
#pragma unroll
                for (int k = 0; k < TW; ++k) {
                    const int col = max(c0 + k - pad, max(c0 - 1, 0));  // 0-based ref column of register k
                    const uint32_t vk = INBAND ? (H[k] >> 2) & 0x3FFF3FFFu : H[k];  // (values are non-negative)
                    const uint32_t cand = add2(vk, pk(fc.gR * (col + 1)));
                    const uint32_t low = 0xFFFFu - (uint32_t)col;
                    key_a = max(key_a, (int)((cand << 16) | low));
                    key_b = max(key_b, (int)((cand & 0xFFFF0000u) | low));
                }
                uint32_t *hk = b.hrow + ((size_t)s * g.duos + duo) * 2;
                if (!SOLO || fw.lane == 0) hk[0] = (uint32_t)key_a;
                if (!SOLO || fw.lane == 1) hk[1] = (uint32_t)key_b;
            } else {  // whole last row (SSEKernel.cpp:1302-1310), un-shifted; column 0 is 0 and `best` starts at 0
#pragma unroll
                for (int k = 0; k < TW; ++k)
                    best = __viaddmax_s16x2(H[k], pk(fc.gF * m + fc.gR * (max(c0 + k - pad, c0 - 1) + 1)), best);
            }
        }
        if (!ALIGN) {
            if (!SOLO || fw.lane == 0) b.scores[b.pair_of[slot_a]] = (int16_t)(best & 0xFFFF);
            if (!SOLO || fw.lane == 1) b.scores[b.pair_of[slot_b]] = (int16_t)(best >> 16);
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

template <bool ALIGN, int TW, int POLICY, bool INBAND = false>
void launch_one(const ChunkGeom &g, const ChunkBuffers &b, const FastConsts &fc, cudaStream_t stream) {
    const int duos = (g.n + 1) / 2;
    if (g.n >= 2) {
        const int threads = Block<ALIGN, TW, false>::NT;
        fill_nw_kernel<ALIGN, TW, false, POLICY, INBAND><<<(duos + threads - 1) / threads, threads, 0, stream>>>(g, b, fc);
    }
    // leftovers of the bucketing: a fixed grid strides over the list the prep kernel compiled (empty on a
    // uniform batch: the blocks read the count and leave)
    if (g.solo) {
        const int threads = Block<ALIGN, TW, true>::NT;
        fill_nw_kernel<ALIGN, TW, true, POLICY, INBAND><<<std::min(2 * ((duos + threads - 1) / threads), 148 * 4), threads, 0, stream>>>(g, b, fc);
    }
}

}  // namespace

int launch_fill_nw(const ChunkGeom &g, const ChunkBuffers &b, int mode, const FastConsts &fc, cudaStream_t stream) {
    const bool align = mode == MODE_NW_ALIGN;
    const bool simd = align && g.policy == 1;
    const bool tagged = align && g.inband != 0;
    if (g.fast_tw == 30) {
        if (tagged && simd) launch_one<true, 30, 1, true>(g, b, fc, stream);
        else if (tagged) launch_one<true, 30, 0, true>(g, b, fc, stream);
        else if (simd) launch_one<true, 30, 1>(g, b, fc, stream);
        else if (align) launch_one<true, 30, 0>(g, b, fc, stream);
        else launch_one<false, 30, 0>(g, b, fc, stream);
    } else {
        if (tagged && simd) launch_one<true, 32, 1, true>(g, b, fc, stream);
        else if (tagged) launch_one<true, 32, 0, true>(g, b, fc, stream);
        else if (simd) launch_one<true, 32, 1>(g, b, fc, stream);
        else if (align) launch_one<true, 32, 0>(g, b, fc, stream);
        else launch_one<false, 32, 0>(g, b, fc, stream);
    }
    return (g.n >= 2 ? 1 : 0) + (g.solo ? 1 : 0);
}

}  // namespace va
