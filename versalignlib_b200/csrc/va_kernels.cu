// va_kernels.cu -- sm_100a kernels of the general (any input, any policy) path and their
// launchers.  The packed 16-bit fast path lives in va_fast.cuh and is launched from here too.
//
// What each kernel stands in for in the reference:
//   prep_kernel        the per-cell char_to_score[] look-ups (DefaultKernel.h:43-60) and the
//                      first-invalid-character scans (DefaultKernel.cpp:308-310,348-350;
//                      SSEKernel.cpp:514-518,673-677), hoisted out of the DP loop
//   fill_general       score_alignment_* and calculate_alignment_matrix_* of every reference
//                      kernel (DefaultKernel.cpp:83-389; SSEKernel.cpp:226-727,1007-1315;
//                      scoring_kernels.cl, alignment_kernels.cl:38-135,239-364)
// (the traceback kernel is in va_traceback.cu)
#include "va_internal.h"
#include "va_device.cuh"
#include "va_fast.cuh"

namespace va {

// ------------------------------------------------------------------------------------------
// prep: raw bytes -> base codes (slot-interleaved uint4 chunks) + per-pair extents
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ int base_code(unsigned c) {
    c &= 0xDFu;  // fold case; bytes >= 0x80 keep bit 7 and fall through to OTHER
    return c == 'A' ? CODE_A : c == 'C' ? CODE_C : c == 'G' ? CODE_G : c == 'T' ? CODE_T : c == 'N' ? CODE_N : CODE_OTHER;
}

struct SeqScan {
    int last_acgt;     // index of the last ACGT base, -1 if none
    int first_other;   // first byte that is neither ACGT nor N (Default/OpenCL "invalid"), L if none
    int first_nonacgt; // first byte that is not ACGT (SSE/AVX "invalid"), L if none
    int n_acgt;
};

// One thread walks one sequence.  16 bases per uint4, written to [chunk][slot].
__device__ __forceinline__ SeqScan encode_sequence(const uint8_t *__restrict__ raw, int L, int chunks, uint4 *__restrict__ out,
                                                   int slots, int slot, bool live) {
    SeqScan s;
    s.last_acgt = -1;
    s.first_other = L;
    s.first_nonacgt = L;
    s.n_acgt = 0;
    for (int c = 0; c < chunks; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int pos = c * 16 + q * 4 + b;
                int code = CODE_OTHER;
                if (live && pos < L) {
                    code = base_code(raw[pos]);
                    if (code < 4) {
                        s.last_acgt = pos;
                        s.n_acgt++;
                    } else {
                        if (pos < s.first_nonacgt) s.first_nonacgt = pos;
                        if (code == CODE_OTHER && pos < s.first_other) s.first_other = pos;
                    }
                }
                word |= (uint32_t)code << (8 * b);
            }
            w[q] = word;
        }
        out[(size_t)c * slots + slot] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return s;
}

__global__ void __launch_bounds__(128) prep_kernel(ChunkGeom g, ChunkBuffers b, int mode, int policy, int trim) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= g.slots) return;
    const bool live = slot < g.n;
    SeqScan rd = encode_sequence(b.raw_reads + (size_t)slot * g.read_length, g.read_length, g.read_chunks, b.code_reads,
                                 g.slots, slot, live);
    SeqScan rf = encode_sequence(b.raw_refs + (size_t)slot * g.ref_length, g.ref_length, g.ref_chunks, b.code_refs,
                                 g.slots, slot, live);
    PairMeta m;
    m.true_rows = (int16_t)(rd.last_acgt + 1);
    m.true_cols = (int16_t)(rf.last_acgt + 1);
    m.flags = (int16_t)((rd.n_acgt != rd.last_acgt + 1 ? 1 : 0) | (rf.n_acgt != rf.last_acgt + 1 ? 2 : 0));
    const int inv_r = policy == 1 ? rd.first_nonacgt : rd.first_other;
    const int inv_f = policy == 1 ? rf.first_nonacgt : rf.first_other;
    m.max_read_pos = (int16_t)(inv_r - 1);
    m.max_ref_pos = (int16_t)(inv_f - 1);
    m.pad = 0;
    if (!live) {
        m.rows = m.cols = 0;
    } else if (mode == MODE_NW_ALIGN) {
        // rows below the first invalid read character are never consulted; the end-cell rule
        // scans the whole padded width of the last valid row
        m.rows = (int16_t)(m.max_read_pos + 1);
        m.cols = (int16_t)g.ref_length;
    } else if (trim) {
        // trailing rows/columns that can only score 0 never change the result while both gap
        // scores are <= 0 (SURVEY.md A.1/A.2 "padding is neutral")
        m.rows = m.true_rows;
        m.cols = m.true_cols;
        if (mode == MODE_SW_ALIGN) {
            // with a zero score traceback starts at cell (0,0) (DefaultKernel.cpp:207-208), so that
            // cell's pointer must exist even when a sequence holds no ACGT base at all
            m.rows = (int16_t)max((int)m.rows, min(1, g.read_length));
            m.cols = (int16_t)max((int)m.cols, min(1, g.ref_length));
        }
    } else {
        m.rows = (int16_t)g.read_length;
        m.cols = (int16_t)g.ref_length;
    }
    b.meta[slot] = m;
}

// ------------------------------------------------------------------------------------------
// general fill: one thread per pair, 32-bit lanes, 16-column register strip
// ------------------------------------------------------------------------------------------

constexpr int GEN_TW = 16;

template <int MODE, int POLICY>
__global__ void __launch_bounds__(128) fill_general_kernel(ChunkGeom g, ChunkBuffers b, Scoring sc) {
    constexpr bool SW = MODE == MODE_SW_SCORE || MODE == MODE_SW_ALIGN;
    constexpr bool ALIGN = MODE == MODE_SW_ALIGN || MODE == MODE_NW_ALIGN;

    __shared__ int sub[64];  // sub[read_code * 8 + ref_code]
    if (threadIdx.x < 64) {
        const int r = threadIdx.x >> 3, f = threadIdx.x & 7;
        sub[threadIdx.x] = (r < 4 && f < 4) ? (r == f ? sc.match : sc.mismatch) : 0;
    }
    __syncthreads();

    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cells = 0;
    // pairs the packed kernel owns (va_fast.cuh) are skipped here
    if (slot < g.n && !duo_is_fast(g, MODE, slot & ~1, b.meta[slot & ~1], b.meta[slot | 1])) {
        const PairMeta meta = b.meta[slot];
        const int m = meta.rows, n = meta.cols;
        const int gF = sc.gap_ref, gR = sc.gap_read;
        cells = (unsigned long long)m * (unsigned long long)n;

        int best = 0, best_i = 0, best_j = 0;  // SW: first strictly greater cell in row-major order
        int border = 0;                        // NW score: max(0, last column, last row)
        int row_max = m * gF, row_idx = 0;     // NW align: arg-max of the last valid row, column 0 first
        const uint32_t *rcodes = reinterpret_cast<const uint32_t *>(b.code_reads);

        for (int c0 = 0; c0 < n; c0 += GEN_TW) {
            const uint4 fq = b.code_refs[(size_t)(c0 >> 4) * g.slots + slot];
            const uint32_t fw[4] = {fq.x, fq.y, fq.z, fq.w};
            int H[GEN_TW];
#pragma unroll
            for (int k = 0; k < GEN_TW; ++k) H[k] = 0;  // row 0 of the matrix is 0 in every mode
            int diag_in = 0;                             // H[i][c0]: value left of the strip, previous row
            const bool last_strip = c0 + GEN_TW >= n;
            uint32_t rword = 0;
            for (int i = 0; i < m; ++i) {
                if ((i & 3) == 0) rword = rcodes[((size_t)(i >> 4) * g.slots + slot) * 4 + ((i >> 2) & 3)];
                const int rc = (rword >> (8 * (i & 3))) & 0xFF;
                const int *srow = sub + rc * 8;
                int left;
                if (c0 == 0) left = MODE == MODE_NW_ALIGN ? (i + 1) * gF : 0;  // matrix column 0
                else left = b.boundary[(size_t)i * g.slots + slot];
                int diag = diag_in;
                diag_in = left;
                uint32_t dirbits = 0;
#pragma unroll
                for (int k = 0; k < GEN_TW; ++k) {
                    const int fc = (fw[k >> 2] >> (8 * (k & 3))) & 0xFF;
                    const int up = H[k];
                    const int d = diag + srow[fc];
                    const int u = up + gF;
                    const int l = left + gR;
                    int h = max(d, max(u, l));
                    if (SW) h = max(h, 0);
                    if (ALIGN) {
                        int code;
                        if (POLICY == 0) {
                            // START (SW and zero) > DIAG > UP > LEFT   (DefaultKernel.cpp:238-248,338-346)
                            code = h == d ? DIR_DIAG : (h == u ? DIR_UP : DIR_LEFT);
                            if (SW && h == 0) code = DIR_START;
                        } else {
                            // max of the codes, DIAG only between two ACGT bases, no zero rule
                            // (SSEKernel.cpp:366-379,646-659)
                            code = DIR_START;
                            if (h == u) code = DIR_UP;
                            if (h == l) code = DIR_LEFT;
                            if (h == d && rc < 4 && fc < 4) code = DIR_DIAG;
                        }
                        dirbits |= (uint32_t)code << (2 * k);
                    }
                    const bool in_range = c0 + k < n;
                    if (MODE == MODE_SW_SCORE) {
                        if (in_range) best = max(best, h);
                    } else if (MODE == MODE_SW_ALIGN) {
                        if (in_range && (h > best || (h == best && i < best_i))) {
                            best = h;
                            best_i = i;
                            best_j = c0 + k;
                        }
                    } else if (MODE == MODE_NW_SCORE) {
                        if (in_range && (c0 + k == n - 1 || i == m - 1)) border = max(border, h);
                    } else {
                        if (in_range && i == m - 1 && h > row_max) {
                            row_max = h;
                            row_idx = c0 + k;
                        }
                    }
                    diag = up;
                    H[k] = h;
                    left = h;
                }
                if (!last_strip) b.boundary[(size_t)i * g.slots + slot] = left;
                if (ALIGN) {
                    const int seg = c0 >> 3;
                    b.dirs[((size_t)seg * g.rows_alloc + i) * g.slots + slot] = (uint16_t)(dirbits & 0xFFFF);
                    if (seg + 1 < g.segs)
                        b.dirs[((size_t)(seg + 1) * g.rows_alloc + i) * g.slots + slot] = (uint16_t)(dirbits >> 16);
                }
            }
        }
        if (MODE == MODE_SW_SCORE) {
            b.scores[slot] = (int16_t)best;
        } else if (MODE == MODE_NW_SCORE) {
            b.scores[slot] = (int16_t)border;
        } else if (MODE == MODE_SW_ALIGN) {
            b.end_cell[2 * slot] = (int16_t)best_i;
            b.end_cell[2 * slot + 1] = (int16_t)best_j;
            b.scores[slot] = (int16_t)best;
        } else {
            // DefaultKernel.cpp:381-387: (max_read_pos, min(max_ref_pos, arg-max of that row))
            b.end_cell[2 * slot] = (int16_t)(m - 1);
            b.end_cell[2 * slot + 1] = (int16_t)min((int)meta.max_ref_pos, row_idx);
            b.scores[slot] = (int16_t)row_max;
        }
    }
    // one atomic per warp for the cell counter
    for (int o = 16; o > 0; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
    if ((threadIdx.x & 31) == 0 && cells) atomicAdd(b.cell_count, cells);
}

// ------------------------------------------------------------------------------------------
// integer-pipe peak: dependent VIADDMNMX / VIMNMX3 chains, registers only
// ------------------------------------------------------------------------------------------

template <int KIND>
__global__ void __launch_bounds__(256) int_peak_kernel(int iters, unsigned int seed, unsigned int *sink) {
    constexpr int CH = 8;
    uint32_t v[CH];
    const uint32_t g = 0xFFFDFFFDu ^ (seed & 1);  // (-3,-3)
    uint32_t w = seed * 2654435761u + threadIdx.x;
#pragma unroll
    for (int c = 0; c < CH; ++c) v[c] = (threadIdx.x + c * 7 + seed) & 0x00FF00FF;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (KIND == 0) v[c] = (uint32_t)__viaddmax_s32((int)v[c], (int)g, (int)w);
                else if (KIND == 1) v[c] = __viaddmax_s16x2(v[c], g, w);
                else if (KIND == 2) v[c] = __viaddmax_s16x2_relu(v[c], g, w);
                else v[c] = __vimax3_s16x2(v[c], g, w);
            }
            w += 0x00010001u;
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= v[c];
    if (acc == 0x12345678u) sink[0] = acc;  // practically never: keeps the chains alive
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------

int launch_prep(const ChunkGeom &g, const ChunkBuffers &b, int mode, int policy, const Scoring &sc, cudaStream_t stream) {
    const int trim = (sc.gap_read <= 0 && sc.gap_ref <= 0) ? 1 : 0;
    const int threads = 128, blocks = (g.slots + threads - 1) / threads;
    prep_kernel<<<blocks, threads, 0, stream>>>(g, b, mode, policy, trim);
    return 1;
}

template <int MODE>
static void launch_fill_mode(const ChunkGeom &g, const ChunkBuffers &b, int policy, const Scoring &sc, cudaStream_t stream) {
    const int threads = 128, blocks = (g.n + threads - 1) / threads;
    if (policy == 0) fill_general_kernel<MODE, 0><<<blocks, threads, 0, stream>>>(g, b, sc);
    else fill_general_kernel<MODE, 1><<<blocks, threads, 0, stream>>>(g, b, sc);
}

int launch_fill_general(const ChunkGeom &g, const ChunkBuffers &b, int mode, int policy, const Scoring &sc,
                        cudaStream_t stream) {
    if (g.n <= 0) return 0;
    switch (mode) {
        case MODE_SW_SCORE: launch_fill_mode<MODE_SW_SCORE>(g, b, 0, sc, stream); break;
        case MODE_NW_SCORE: launch_fill_mode<MODE_NW_SCORE>(g, b, 0, sc, stream); break;
        case MODE_SW_ALIGN: launch_fill_mode<MODE_SW_ALIGN>(g, b, policy, sc, stream); break;
        default: launch_fill_mode<MODE_NW_ALIGN>(g, b, policy, sc, stream); break;
    }
    return 1;
}

int launch_int_peak(int kind, int sm_count, int iters, unsigned int *sink, cudaStream_t stream, double *lane_ops) {
    const int threads = 256, blocks = sm_count * 8;
    switch (kind) {
        case 0: int_peak_kernel<0><<<blocks, threads, 0, stream>>>(iters, 1u, sink); break;
        case 1: int_peak_kernel<1><<<blocks, threads, 0, stream>>>(iters, 1u, sink); break;
        case 2: int_peak_kernel<2><<<blocks, threads, 0, stream>>>(iters, 1u, sink); break;
        default: int_peak_kernel<3><<<blocks, threads, 0, stream>>>(iters, 1u, sink); break;
    }
    const double lanes = kind == 0 ? 1.0 : 2.0;
    *lane_ops = (double)blocks * threads * (double)iters * 8.0 * 8.0 * lanes;
    return 1;
}

}  // namespace va
