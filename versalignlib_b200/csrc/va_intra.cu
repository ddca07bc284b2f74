// va_intra.cu -- intra-task Smith-Waterman score kernel for long pairs: ONE WARP computes one
// pair-of-pairs (two pairs in the two s16 lanes of every register), walking the matrix as a
// skewed wavefront.  The reference has nothing like it -- all of its kernels walk one matrix
// row-major with a loop-carried dependency through memory -- so parity is defined
// by the recurrence alone (DefaultKernel.cpp:83-138): the cell arithmetic is the packed kernel's
// (va_fast.cu), only the order of evaluation changes.
//
//   * lane l owns TW consecutive ref columns of the current pass (32*TW columns per pass);
//   * at step t lane l computes row t-l of its columns, so its left neighbour's right edge
//     (computed one step earlier) arrives by __shfl_up_sync together with that row's two
//     substitution tables; lane 0 takes its inputs from a per-32-steps batch that all lanes
//     load coalesced (read codes -> tables, previous pass's boundary column);
//   * lane 31's right edge is collected over 32 steps and stored coalesced as the next pass's
//     boundary column;
//   * the passes of one duo are spread over the warps of a CTA and run as a pipeline (see the kernel).
// Used when a batch has too few pairs to fill the GPU with one thread per pair-of-pairs and
// the pairs are long (launch_fill_intra decides).
#include "va_fast.cuh"

namespace va {

namespace {

constexpr uint32_t NEG2 = 0x80008000u;
__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { return __viaddmax_s16x2(a, b, NEG2); }

constexpr int MAX_PASSES = 64;  // 32 000 columns / (32 lanes x 16 columns)

// One CTA = one duo; its W warps take the column passes round-robin (warp w: passes w, w+W, ...) and run
// them as a pipeline: pass p+1 follows pass p a few dozen rows behind, reading the boundary column pass p
// leaves in global memory (32 rows per coalesced store) as soon as a shared-memory progress counter says
// those rows are there.  A batch of a few long pairs then fills the GPU with W times as many warps.
template <int TW>
__global__ void __launch_bounds__(128) fill_intra_sw_score_kernel(ChunkGeom g, ChunkBuffers b, FastConsts fc) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint32_t T[8];
    __shared__ int prog[MAX_PASSES];      // rows of pass p whose right edge is in global memory
    __shared__ uint32_t warp_best[4];
    if (threadIdx.x < 8) T[threadIdx.x] = fc.tab[threadIdx.x];
    for (int t = threadIdx.x; t < MAX_PASSES; t += blockDim.x) prog[t] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int duo = blockIdx.x;
    const int slot_a = 2 * duo, slot_b = slot_a + 1;
    if (slot_b >= g.n) return;  // CTA-uniform
    const PairMeta ma = b.meta[slot_a], mb = b.meta[slot_b];
    if (!duo_is_fast(g, MODE_SW_SCORE, slot_a, ma, mb)) return;
    const int m = max((int)ma.rows, (int)mb.rows), n = ma.cols;
    const uint32_t gF2 = fc.gF2, gR2 = fc.gR2, ngR2 = fc.dFR2;  // dFR2 = -gap_read in the score modes
    const uint8_t *cread = reinterpret_cast<const uint8_t *>(b.code_reads) + (size_t)slot_a * 16;
    const uint8_t *cref = reinterpret_cast<const uint8_t *>(b.code_refs);
    const uint32_t chunk_stride = (uint32_t)g.slots * 16u;
    uint32_t *bnd = b.fboundary + (size_t)duo * g.rows_alloc;  // this duo's boundary column (rewritten pass after pass)
    volatile int *vprog = prog;

    uint32_t best = 0;
    const int pass_cols = 32 * TW;
    const int npasses = (n + pass_cols - 1) / pass_cols;
    const int steps = m + 31;
    for (int pass = warp; pass < npasses; pass += W) {
        const int c_base = pass * pass_cols;
        const bool first_pass = pass == 0, last_pass = pass == npasses - 1;
        const int c0 = c_base + lane * TW;
        const int kv = min(TW, max(0, n - c0));  // my valid columns in this pass
        uint32_t sel[TW], H[TW];
#pragma unroll
        for (int k = 0; k < TW; ++k) {
            const int col = min(c0 + k, n - 1);
            const size_t off = (size_t)(col >> 4) * chunk_stride + (col & 15);
            const uint32_t fa = cref[off + (size_t)slot_a * 16], fb = cref[off + (size_t)slot_b * 16];
            sel[k] = fa | ((fa | 8u) << 4) | ((fb | 4u) << 8) | ((fb | 12u) << 12);
            H[k] = 0u;
        }
        uint32_t diag_next = 0u;                        // H[row][c0] of the previous row, 0 for matrix row 0
        uint32_t cur_ta = 0, cur_tb = 0, edge = gR2;    // what this lane used / produced at its last step
        uint32_t out_keep = 0;                          // lane 31's right edges, one per lane, for the coalesced store

        for (int t0 = 0; t0 < steps; t0 += 32) {
            // batch inputs of lane 0 for steps t0..t0+31: row r = t0 + lane
            const int r = t0 + lane;
            uint32_t bat_a = 0, bat_b = 0, bat_left = gR2;  // matrix column 0 is 0: "left + gR" = gR
            if (!first_pass && t0 < m) {
                // the previous pass (another warp of this CTA when W > 1) must have left these rows
                const int need = min(t0 + 32, m);
                while (vprog[pass - 1] < need) {
                }
                __threadfence_block();
            }
            if (r < m) {
                const uint32_t roff = (uint32_t)(r >> 4) * chunk_stride + (uint32_t)(r & 15);
                bat_a = T[cread[roff]];
                bat_b = T[cread[roff + 16]];
                if (!first_pass) bat_left = __ldcg(bnd + r);  // written by another warp: read it where it was written (L2)
            }
            __syncwarp();  // every lane has read its row before this warp overwrites the column below
            const int s_end = min(32, steps - t0);
            for (int s = 0; s < s_end; ++s) {
                const int t = t0 + s;
                // lane 0 reads the batch, every other lane takes what its left neighbour used last step
                const uint32_t a0 = __shfl_sync(FULL, bat_a, s), b0 = __shfl_sync(FULL, bat_b, s), l0 = __shfl_sync(FULL, bat_left, s);
                const uint32_t pa = __shfl_up_sync(FULL, cur_ta, 1), pb = __shfl_up_sync(FULL, cur_tb, 1), pl = __shfl_up_sync(FULL, edge, 1);
                const uint32_t ta = lane == 0 ? a0 : pa, tb = lane == 0 ? b0 : pb;
                uint32_t left = lane == 0 ? l0 : pl;
                const int row = t - lane;
                if (row >= 0 && row < m) {
                    cur_ta = ta;
                    cur_tb = tb;
                    uint32_t diag = diag_next;
                    diag_next = add2(left, ngR2);
                    if (kv == TW) {
#pragma unroll
                        for (int k = 0; k < TW; ++k) {
                            const uint32_t sub = prmt(ta, tb, sel[k]);
                            const uint32_t up = H[k];
                            const uint32_t tt = __viaddmax_s16x2(up, gF2, left);
                            const uint32_t h = __viaddmax_s16x2_relu(diag, sub, tt);
                            left = add2(h, gR2);
                            H[k] = h;
                            if (k & 1) best = __vimax3_s16x2(best, h, H[k - 1]);
                            diag = up;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < TW; ++k) {
                            const uint32_t sub = prmt(ta, tb, sel[k]);
                            const uint32_t up = H[k];
                            const uint32_t tt = __viaddmax_s16x2(up, gF2, left);
                            const uint32_t h = __viaddmax_s16x2_relu(diag, sub, tt);
                            left = add2(h, gR2);
                            H[k] = h;
                            if (k < kv) best = __vmaxs2(best, h);  // columns past n stay out of the maximum
                            diag = up;
                        }
                    }
                    edge = left;
                }
                // lane 31 finished row t-31: park its right edge in lane (t-31)&31 until 32 are there
                if (!last_pass) {
                    const uint32_t v = __shfl_sync(FULL, edge, 31);
                    const int orow = t - 31;
                    if (orow >= 0) {
                        if (lane == (orow & 31)) out_keep = v;
                        if ((orow & 31) == 31 || orow == m - 1) {
                            const int row0 = orow & ~31;
                            if (row0 + lane <= orow) bnd[row0 + lane] = out_keep;
                            __threadfence_block();
                            __syncwarp();
                            if (lane == 0) vprog[pass] = orow + 1;  // rows [0, orow] of this pass are out
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = __vmaxs2(best, __shfl_xor_sync(FULL, best, o));
    if (lane == 0) warp_best[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < W; ++w) best = __vmaxs2(best, warp_best[w]);
        b.scores[b.pair_of[slot_a]] = (int16_t)(best & 0xFFFF);
        b.scores[b.pair_of[slot_b]] = (int16_t)(best >> 16);
        atomicAdd(b.cell_count, ((unsigned long long)ma.rows + (unsigned long long)mb.rows) * (unsigned long long)n);
    }
}

}  // namespace

// The intra-task kernel pays when one-thread-per-duo cannot fill the machine: few, long pairs.
bool intra_preferred(int mode, int n_pairs, int read_length, int ref_length, int sm_count) {
    if (mode != MODE_SW_SCORE) return false;
    const long long duos = n_pairs / 2;
    return ref_length >= 1024 && read_length >= 256 && duos < (long long)sm_count * 512;
}

int launch_fill_intra(const ChunkGeom &g, const ChunkBuffers &b, int mode, const FastConsts &fc, cudaStream_t stream) {
    if (mode != MODE_SW_SCORE || g.n < 2) return 0;
    const int duos = g.n / 2;
    // warps per duo: as many as there are column passes to pipeline, at most 4 (8 measured no better)
    const int passes = (g.ref_length + 32 * 16 - 1) / (32 * 16);
    const int warps = passes >= 4 ? 4 : passes >= 2 ? 2 : 1;
    fill_intra_sw_score_kernel<16><<<duos, 32 * warps, 0, stream>>>(g, b, fc);
    return 1;
}

}  // namespace va
