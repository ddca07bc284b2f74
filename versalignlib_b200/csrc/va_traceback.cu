// va_traceback.cu -- per-pair traceback over the 2-bit direction matrix and output of the two
// gapped strings, right-aligned in read_length+ref_length byte blocks like the reference's
// calc_alignment_* (DefaultKernel.cpp:391-525; SSEKernel.cpp:729-1005;
// alignment_kernels.cl:146-192,370-414).
//
// Two phases, so that the only dependent chain through HBM is the walk:
//   walk  one thread per pair: follow the pointers from the end cell to START (L2/HBM latency hidden by
//         occupancy); the moves go into a per-thread 2-bit queue in shared memory (global memory when the
//         sequences are too long for that).  The walk over the packed kernels' bit planes is written for
//         instruction count -- the kernel is issue bound, not latency bound: one 8-byte load per matrix
//         row (the row's DIAG and UP planes of 16 columns x both lanes), moves derived from the two bits
//         without branches, the slow path (leaving a 16-column group / a strip) out of line;
//   emit  one warp per pair, 32 moves per step: replay the queue, fetch the read/ref bytes it names
//         and write both strings backwards -- consecutive lanes, consecutive bytes.
// Three result containers:
//   strings, fixed stride    aln_read / aln_ref [pair][L], right aligned (the reference's layout)
//   strings, compact         (host pipeline of the legacy boundary) each pair's two strings back to back,
//                            NUL terminated, at an offset handed out by one atomic per block: half the
//                            D2H bytes of the fixed-stride layout and nothing for the host to skip over
//   moves                    (packed entry points) CIGAR runs / raw 2-bit moves + sequence coordinates
// The packed NW fill kernel leaves the end-cell decision (arg-max of the last valid row,
// DefaultKernel.cpp:352-355,381-387) to this kernel: `hrow` holds that row's per-strip keys.
#include <climits>

#include <cub/device/device_scan.cuh>

#include "va_fast.cuh"

namespace va {

namespace {

constexpr int TB_THREADS = 128;

// Backwards byte reader over one raw sequence: keeps the aligned 16 bytes around the last position in
// registers, so a walk that visits every base costs one 16-byte load per 16 bases instead of one
// sector-sized L2 request per base (with full occupancy the L1 is too small to hold every thread's
// window).  The window never reaches past `limit` (end of the whole raw buffer): there it falls back
// to byte loads.
struct ByteWindow {
    const uint8_t *base;
    const uint8_t *limit;
    uintptr_t have;
    uint4 w;
    __device__ __forceinline__ ByteWindow(const uint8_t *b, const uint8_t *lim) : base(b), limit(lim), have(~(uintptr_t)0) {
        w = make_uint4(0, 0, 0, 0);
    }
    __device__ __forceinline__ unsigned get(int idx) {
        const uint8_t *p = base + idx;
        const uintptr_t a16 = reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)15;
        if (a16 + 16 > reinterpret_cast<uintptr_t>(limit)) return *p;
        if (a16 != have) {
            w = *reinterpret_cast<const uint4 *>(a16);
            have = a16;
        }
        const unsigned k = (unsigned)(reinterpret_cast<uintptr_t>(p) & 15);
        const uint32_t word = k < 8 ? (k < 4 ? w.x : w.y) : (k < 12 ? w.z : w.w);
        return (word >> (8 * (k & 3))) & 0xFFu;
    }
};

// 2-bit move queue of one thread: word w lives at q[w * stride]
struct MoveQueue {
    uint32_t *q;
    size_t stride;
};

template <bool NW>
__global__ void __launch_bounds__(TB_THREADS) traceback_kernel(ChunkGeom g, ChunkBuffers b, Scoring sc, uint32_t *gq,
                                                               int queue_words) {
    const int gap_ref = sc.gap_ref;
    extern __shared__ uint32_t sq[];
    __shared__ int s_moves[TB_THREADS], s_pair[TB_THREADS];
    __shared__ const uint8_t *s_rd[TB_THREADS], *s_rf[TB_THREADS];  // the bases the first move of each pair reads
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_warp_sum[TB_THREADS / 32];
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const bool moves_only = b.moves_out != nullptr;
    const bool compact = b.aln_compact != nullptr;
    const int L = g.read_length + g.ref_length;
    s_moves[threadIdx.x] = -1;
    int my_moves = -1;
    if (slot < g.n) {
    MoveQueue mq;
    if (moves_only) {  // set below: the result region of the pair
        mq.q = nullptr;
        mq.stride = 1;
    } else if (gq) {
        mq.q = gq + slot;
        mq.stride = (size_t)g.slots;
    } else {
        mq.q = sq + threadIdx.x;
        mq.stride = TB_THREADS;
    }
    const PairMeta meta = b.meta[slot];
    const int rows = meta.rows, cols = meta.cols;
    const int mode = NW ? MODE_NW_ALIGN : MODE_SW_ALIGN;
    const PairMeta meta_other = b.meta[slot ^ 1];
    const int owner = slot_owner(g, mode, slot, (slot & 1) ? meta_other : meta, (slot & 1) ? meta : meta_other);
    const bool packed = owner != OWN_NONE;
    // packed NW align end-aligns the lanes of a duo: this lane's matrix row r is sweep row r + row_off
    const int row_off = (NW && owner == OWN_DUO) ? nw_row_offset(g, meta, meta_other) : 0;
    const int duo = slot >> 1, lane_shift = (slot & 1) * 16;
    const int pair = b.pair_of[slot];  // raw bytes and results are indexed in the caller's pair order

    // ---- end cell ----------------------------------------------------------------------
    int i, j;
    if (packed && NW) {
        // First strictly greater column of matrix row `rows`, starting from column 0 = rows*gap_ref.  The fill
        // kernel left, per strip, the best of its columns as one key per lane: (value << 16) | (0xFFFF - column)
        // with value = H - gap_ref*rows, so the signed maximum over the strips (and over column 0's own key) is
        // the greatest value and, among equals, the smallest column (va_nw.cu).
        int key = 0x0000FFFF;  // matrix column 0: value 0, reported as column 0
        {
            const int nstrips = (cols + g.fast_tw - 1) / g.fast_tw;
            const uint32_t *hk = b.hrow + (size_t)duo * 2 + (slot & 1);
            for (int st = 0; st < nstrips; ++st) key = max(key, (int)hk[(size_t)st * g.duos * 2]);
        }
        const int best = (key >> 16) + rows * gap_ref, idx = 0xFFFF - (key & 0xFFFF);
        b.scores[pair] = (int16_t)best;
        i = rows - 1;
        // Pad columns (past `cols`, never filled) take part in the arg-max of the reference.  With both
        // gaps <= 0 one of them beats the best true cell exactly when a value of the last true column in
        // one of the min(pad columns, rows) matrix rows above comes down a zero-score diagonal; the
        // arg-max is then past max_ref_pos and gets clipped to it (DefaultKernel.cpp:387).
        const int pad_cols = g.ref_length - cols, reach = min(pad_cols, rows);
        const int vshift = g.inband ? 2 : 0;  // the tagged fill kernel keeps 4V in the boundary column (va_nw.cu)
        int col_max = (reach == rows && rows > 0) ? 0 : INT_MIN;  // matrix row 0
        if (reach > 0) {
            const uint32_t *bl = b.fboundary + duo;  // last true column (matrix column `cols`), shifted
            for (int r = rows - 2; r >= rows - 1 - reach && r >= 0; --r)
                col_max = max(col_max, ((int)(int16_t)(bl[(size_t)(r + row_off) * g.duos] >> lane_shift) >> vshift) + (r + 1) * gap_ref + cols * sc.gap_read);
        }
        j = (pad_cols > 0 && col_max > best) ? (int)meta.max_ref_pos : min((int)meta.max_ref_pos, idx);
        b.end_cell[2 * pair] = (int16_t)i;
        b.end_cell[2 * pair + 1] = (int16_t)j;
    } else {
        i = b.end_cell[2 * pair];
        j = b.end_cell[2 * pair + 1];
    }
    const int end_i = i, end_j = j;

    // ---- walk ----------------------------------------------------------------------------
    // One walk from the end cell to START; every move goes to `sink(code, t)` (t = moves so far).
    // final_i / final_j: the cell the walk stopped in (0-based sequence coordinates, -1 = matrix row / column 0)
    int final_i = end_i, final_j = end_j;
    auto walk = [&](auto &&sink) -> int {
        int i = end_i, j = end_j;
        int n_moves = 0;
        if (packed) {
            // The packed kernels' direction words: uint4 per (strip, ROW PAIR, group of 16 columns, duo), viewed
            // here as two uint2 halves (even row, odd row): .x = DIAG plane, .y = UP plane, low 16 bits lane A.
            const int tw = g.fast_tw, ng = fast_groups(tw);
            int strip = j >= 0 ? j / tw : 0, k = j >= 0 ? j - strip * tw : 0;
            // NW: a partial last strip keeps its true columns in the LAST registers of the strip (va_nw.cu)
            int kmin = 0;
            if (NW && cols > 0 && strip == (cols - 1) / tw) {
                kmin = tw - (cols - strip * tw);
                k += kmin;
            }
            const bool in_range = i >= 0 && i < rows && j < cols;
            // SW under the SSE/AVX policy (g.policy 1): no zero rule -- the walk follows the pointers through zero cells until
            // the third plane says START (h < 0 before the floor); it begins even when the score is 0
            const bool sw_simd = !NW && g.policy == 1;
            if (in_range && j >= 0 && (NW || sw_simd || (int)b.scores[pair] > 0)) {
                // 32-bit offsets in uint2 units (the host keeps a chunk's direction region below 2^32 of them)
                const uint32_t group_step2 = 2u * (uint32_t)g.duos;                      // next group
                const uint32_t pair_step2 = (uint32_t)ng * group_step2;                  // next row pair
                const uint32_t strip_step2 = (uint32_t)fast_row_pairs(g) * pair_step2;   // next strip
                int si = i + row_off;  // sweep row: where the fill kernel stored this matrix row
                const uint2 *base2 = reinterpret_cast<const uint2 *>(b.fdirs) + 2 * (size_t)duo;
                uint32_t off = (uint32_t)strip * strip_step2 + (uint32_t)(si >> 1) * pair_step2 + (uint32_t)(k >> 4) * group_step2 + (uint32_t)(si & 1);
                const uint32_t up_even = pair_step2 - 1u;  // from an even sweep row to the odd row above it
                const uint32_t strip_back = strip_step2 - (uint32_t)((tw - 1) >> 4) * group_step2;  // to the last group of the strip before
                int bit = lane_shift + (k & 15);
                const uint32_t *z32 = reinterpret_cast<const uint32_t *>(b.fdirs_z) + 2 * (size_t)duo;  // same numbering as base2
                uint2 w = __ldg(base2 + off);
                // SW: the packed fill stores the pointer a cell would have in NW; a cell whose value is 0 is
                // START (DefaultKernel.cpp:240-241).  The value is known along the path: it starts at the best
                // score and every move gives back what it added.
                int hval = NW ? 1 : (int)b.scores[pair];
                const uint32_t pol = (uint32_t)g.policy & 1u;
                ByteWindow wread(seq_ptr(b.raw_reads, b.read_off, pair, g.read_length), seq_end(b.raw_reads, b.read_off, g.n, g.read_length));
                ByteWindow wref(seq_ptr(b.raw_refs, b.ref_off, pair, g.ref_length), seq_end(b.raw_refs, b.ref_off, g.n, g.ref_length));
                // NW, tagged direction words (ChunkGeom::inband, va_nw.cu): .x = columns 0..7 of the group, .y = columns
                // 8..15; column c of a word at bits 2*(7 - c) of the lane's half: 2 = DIAG, 1 = the move that wins the
                // UP / LEFT tie under the call's policy, 0 = the other one
                const bool tagged = NW && g.inband != 0;
                if (tagged) {
                    // Its own loop, written without branches: the 32 walks of a warp leave their groups and strips at
                    // different steps, so a side path taken by one lane is paid for by all of them on nearly every step.
                    // DIR_UP = 1, DIR_LEFT = 2, DIR_DIAG = 3: bit 0 of the code is "one row up", bit 1 "one column left".
                    const uint32_t first = pol ? DIR_LEFT : DIR_UP, second = pol ? DIR_UP : DIR_LEFT;
                    const uint32_t lut = second | (first << 2) | ((uint32_t)DIR_DIAG << 4);
                    const int sh0 = lane_shift + 14;
                    while (true) {
                        const uint32_t word = (k & 8) ? w.y : w.x;
                        const uint32_t tag = (word >> (sh0 - 2 * (k & 7))) & 3u;
                        const uint32_t code = (lut >> (2u * tag)) & 3u;
                        sink((int)code, n_moves);
                        ++n_moves;
                        const uint32_t um = code & 1u, lm = code >> 1;
                        off -= um ? ((si & 1) ? 1u : up_even) : 0u;
                        si -= (int)um;
                        i -= (int)um;
                        const bool xs = lm && k == kmin;             // leaving the strip
                        const bool xg = lm && (k & 15) == 0 && !xs;  // leaving the 16-column group
                        off -= xs ? strip_back + (uint32_t)(k >> 4) * group_step2 : (xg ? group_step2 : 0u);
                        k = xs ? tw - 1 : k - (int)lm;
                        kmin = xs ? 0 : kmin;
                        j -= (int)lm;
                        if ((i | j) < 0) break;
                        w = __ldg(base2 + off);
                    }
                } else
                while (true) {
                    if (sw_simd && !((__ldg(z32 + off) >> bit) & 1u)) break;  // START
                    const uint32_t dbit = (w.x >> bit) & 1u, ubit = (w.y >> bit) & 1u;
                    // second plane: UP >= LEFT (policy 0) or LEFT >= UP (policy 1, SSE/AVX tie order)
                    const int code = dbit ? DIR_DIAG : ((ubit ^ pol) ? DIR_UP : DIR_LEFT);
                    sink(code, n_moves);
                    ++n_moves;
                    if (!NW && !sw_simd) {
                        if (code == DIR_UP) hval -= sc.gap_ref;
                        else if (code == DIR_LEFT) hval -= sc.gap_read;
                        else {
                            const unsigned ca = wread.get(i) & 0xDFu, cb = wref.get(j) & 0xDFu;
                            const bool va = ca == 'A' || ca == 'C' || ca == 'G' || ca == 'T', vb = cb == 'A' || cb == 'C' || cb == 'G' || cb == 'T';
                            hval -= (va && vb) ? (ca == cb ? sc.match : sc.mismatch) : 0;
                        }
                    }
                    if (code != DIR_LEFT) {  // one matrix row up
                        off -= (si & 1) ? 1u : up_even;
                        --si;
                        --i;
                    }
                    if (code != DIR_UP) {  // one column left
                        --j;
                        if (k == kmin) {  // leaving the strip
                            kmin = 0;
                            off -= strip_back + (uint32_t)(k >> 4) * group_step2;
                            k = tw - 1;
                            bit = lane_shift + (k & 15);
                        } else {
                            if ((k & 15) == 0) {  // leaving the 16-column group
                                off -= group_step2;
                                bit += 16;
                            }
                            --k;
                            --bit;
                        }
                    }
                    if ((i | j) < 0 || (!NW && !sw_simd && hval <= 0)) break;
                    w = __ldg(base2 + off);
                }
            }
            // matrix column 0 (DefaultKernel.cpp:304): NW walks up to row 0, SW stops
            if (NW && i >= 0 && i < rows && j < 0) {
                for (; i >= 0; --i) {
                    sink(DIR_UP, n_moves);
                    ++n_moves;
                }
            }
        } else if (g.affine) {
            // affine-gap variant (fill_general_kernel<.., AFFINE>): 4 bits per cell -- where H came from, and whether the E / F
            // gap state of the cell was opened from H; the walk is the three-state machine of oracle/va_oracle_affine.c
            int state = 0;  // 0 = H, 1 = F (a run of UP moves), 2 = E (a run of LEFT moves)
            while (true) {
                int code;
                if (i < 0 || i >= rows || j >= cols) code = DIR_START;
                else if (j < 0) code = NW ? (DIR_UP | (i == 0 ? 8 : 0)) : DIR_START;  // column 0: one leading gap, opened in row 0
                else code = (b.dirs4[((size_t)(j >> 3) * g.rows_alloc + i) * g.slots + slot] >> (4 * (j & 7))) & 15;
                if (state == 0) {
                    const int hp = code & 3;
                    if (hp == DIR_START) break;
                    if (hp == DIR_UP) { state = 1; continue; }
                    if (hp == DIR_LEFT) { state = 2; continue; }
                    sink(DIR_DIAG, n_moves);
                    --i;
                    --j;
                } else if (state == 1) {
                    sink(DIR_UP, n_moves);
                    if (code & 8) state = 0;
                    --i;
                } else {
                    sink(DIR_LEFT, n_moves);
                    if (code & 4) state = 0;
                    --j;
                }
                ++n_moves;
            }
        } else {
            while (true) {
                int code;
                if (i < 0 || i >= rows || j >= cols) code = DIR_START;
                else if (j < 0) code = NW ? DIR_UP : DIR_START;
                else code = (b.dirs[((size_t)(j >> 3) * g.rows_alloc + i) * g.slots + slot] >> (2 * (j & 7))) & 3;
                if (code == DIR_START) break;
                sink(code, n_moves);
                ++n_moves;
                if (code != DIR_LEFT) --i;
                if (code != DIR_UP) --j;
            }
        }
        final_i = i;
        final_j = j;
        return n_moves;
    };
    // 2-bit queue sink: 16 moves per word
    uint32_t acc = 0;
    uint32_t *qcur = mq.q;  // where the word being filled goes
    auto queue_sink = [&](int code, int t) {
        acc |= (uint32_t)code << (2 * (t & 15));
        if ((t & 15) == 15) {
            *qcur = acc;
            qcur += mq.stride;
            acc = 0;
        }
    };
    int n_moves;
    if (moves_only) {
        // The moves ARE the result (packed entry points).  They leave as CIGAR runs in walk order (length << 4 |
        // op, last run of the alignment first) when a pair has at most queue_words of them -- word 0 = run count
        // -- else as the raw 2-bit queue from a second walk, word 0 = 0x80000000 | moves.
        uint32_t *out = b.moves_out + (size_t)pair * (queue_words + 1);
        int cur = -1, len = 0, nruns = 0;
        auto rle_sink = [&](int code, int) {
            const int op = code == DIR_DIAG ? 0 : code == DIR_UP ? 1 : 2;
            if (op == cur) {
                ++len;
            } else {
                if (len) {
                    if (nruns < queue_words) out[1 + nruns] = ((uint32_t)len << 4) | (uint32_t)cur;
                    ++nruns;
                }
                cur = op;
                len = 1;
            }
        };
        n_moves = walk(rle_sink);
        if (len) {
            if (nruns < queue_words) out[1 + nruns] = ((uint32_t)len << 4) | (uint32_t)cur;
            ++nruns;
        }
        if (nruns <= queue_words) {
            out[0] = (uint32_t)nruns;
        } else {
            mq.q = qcur = out + 1;
            mq.stride = 1;
            n_moves = walk(queue_sink);
            if (n_moves & 15) *qcur = acc;
            out[0] = 0x80000000u | (uint32_t)n_moves;
        }
        if (b.coords) {  // aligned region in sequence coordinates, 0-based, half open
            int32_t *co = b.coords + 4 * (size_t)pair;
            co[0] = final_i + 1;
            co[1] = end_i + 1;
            co[2] = final_j + 1;
            co[3] = end_j + 1;
        }
        if (b.run_count) b.run_count[pair] = (uint32_t)nruns;
    } else {
        n_moves = walk(queue_sink);
        if (n_moves & 15) *qcur = acc;
    }

    if (moves_only) {
        b.start[pair] = (int16_t)(L - 1 - n_moves);
    } else {
        my_moves = n_moves;
        s_moves[threadIdx.x] = n_moves;
        s_rd[threadIdx.x] = seq_ptr(b.raw_reads, b.read_off, pair, g.read_length) + end_i;
        s_rf[threadIdx.x] = seq_ptr(b.raw_refs, b.ref_off, pair, g.ref_length) + end_j;
        s_pair[threadIdx.x] = pair;
    }
    }  // slot < g.n
    if (moves_only) return;
    // compact strings: this block's pairs take one contiguous piece of the output, 2 * (moves + 1) bytes each
    __shared__ uint32_t s_off[TB_THREADS];
    if (compact) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t bytes = my_moves >= 0 ? 2u * (uint32_t)(my_moves + 1) : 0u;
        uint32_t incl = bytes;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp_sum[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t total = 0;
            for (int wv = 0; wv < TB_THREADS / 32; ++wv) {
                const uint32_t v = s_warp_sum[wv];
                s_warp_sum[wv] = total;
                total += v;
            }
            s_base = atomicAdd(b.compact_cursor, (unsigned long long)total);
        }
        __syncthreads();
        const uint32_t rel = s_warp_sum[warp] + (incl - bytes);  // offset inside the block's piece
        s_off[threadIdx.x] = rel;
        if (my_moves >= 0) b.compact_off[s_pair[threadIdx.x]] = (uint32_t)(s_base + rel);
        __syncwarp();
    } else {
        __syncthreads();
    }

    // ---- emit ----------------------------------------------------------------------------
    // Warp-cooperative: the walks of the block are done, their move queues sit in shared (or global)
    // memory.  Each warp replays the queues of its 32 threads one pair at a time, 32 moves per step: lane l
    // takes move t0+l, two ballots give every lane how many read / ref bases the moves before it consumed,
    // so consecutive lanes fetch consecutive bases and write consecutive bytes of both gapped strings --
    // every load and store of the step is one or two sectors.
    const int lane = threadIdx.x & 31, warp_first = threadIdx.x & ~31;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int q = warp_first; q < warp_first + 32; ++q) {
        const int n_moves = s_moves[q];
        if (n_moves < 0) continue;  // past the end of the chunk
        const int pair = s_pair[q];
        const uint32_t *qq = gq ? gq + (blockIdx.x * blockDim.x + q) : sq + q;
        const size_t qstride = gq ? (size_t)g.slots : (size_t)TB_THREADS;
        // cursors: the base this step's first move reads / the NUL byte behind each string (move t goes to [-1 - t])
        const uint8_t *rd = s_rd[q], *rf = s_rf[q];
        uint8_t *oa, *ob;
        if (compact) {
            oa = b.aln_compact + (size_t)(s_base + s_off[q]) + n_moves;
            ob = oa + n_moves + 1;
        } else {
            oa = b.aln_read + (size_t)pair * L + (L - 1);
            ob = b.aln_ref + (size_t)pair * L + (L - 1);
        }
        for (int t0 = 0; t0 < n_moves; t0 += 32) {
            const int t = t0 + lane;
            const bool valid = t < n_moves;
            const uint32_t word = valid ? qq[(size_t)(t >> 4) * qstride] : 0u;
            const int code = (word >> (2 * (t & 15))) & 3;
            const bool takes_r = valid && code != DIR_LEFT, takes_f = valid && code != DIR_UP;
            const unsigned mr = __ballot_sync(0xffffffffu, takes_r), mf = __ballot_sync(0xffffffffu, takes_f);
            uint8_t a = '-', c = '-';
            if (takes_r) a = *(rd - __popc(mr & lt_mask));
            if (takes_f) c = *(rf - __popc(mf & lt_mask));
            // fixed stride: a walk longer than L - 1 moves (only when every move is a gap) is cut at the block start
            if (valid && (compact || t <= L - 2)) {
                oa[-1 - t] = a;
                ob[-1 - t] = c;
            }
            rd -= __popc(mr);
            rf -= __popc(mf);
        }
        if (lane == 0) {
            // start may be negative only when every move was a gap (never with gap scores < 0)
            b.start[pair] = (int16_t)(L - 1 - n_moves);
            if (compact || L >= 1) {
                *oa = 0;
                *ob = 0;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// Long pairs (the chunk was filled by the intra-task kernels, va_intra.cu): ONE WARP per pair.
//   walk  the direction words of a duo lie at [duo][16-column strip][row], consecutive rows = consecutive 8-byte
//         words.  The warp keeps a window of 32 rows x 2 strips in registers (lane l: the words of row wr0 + l of
//         strips ws and ws-1, two coalesced 256-byte loads) and follows the path through it with shuffles -- all
//         lanes carry the same walk state -- so HBM latency is paid once per ~24 moves instead of once per move;
//         a one-thread walk of a 10 kbp x 12 kbp pair is ~22 000 dependent loads.
//   emit  the same warp replays the move queue (global memory) 32 moves per step, like traceback_kernel.
// Pairs the general kernel computed (dirty refs, SSE/AVX policy in SW) walk its [segment][row][slot] half-words,
// one broadcast load per move.
// ------------------------------------------------------------------------------------------------------------
template <bool NW>
__global__ void __launch_bounds__(TB_THREADS) traceback_long_kernel(ChunkGeom g, ChunkBuffers b, Scoring sc, uint32_t *gq, int queue_words) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int slot = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (slot >= g.n) return;  // warp-uniform
    const bool moves_only = b.moves_out != nullptr;
    const bool compact = b.aln_compact != nullptr;
    const int L = g.read_length + g.ref_length;
    const int gap_ref = sc.gap_ref;
    const PairMeta meta = b.meta[slot];
    const int rows = meta.rows, cols = meta.cols;
    const int mode = NW ? MODE_NW_ALIGN : MODE_SW_ALIGN;
    const PairMeta meta_other = b.meta[slot ^ 1];
    const bool packed = slot_owner(g, mode, slot, (slot & 1) ? meta_other : meta, (slot & 1) ? meta : meta_other) == OWN_DUO;
    const int duo = slot >> 1, lane_shift = (slot & 1) * 16;
    const int pair = b.pair_of[slot];
    const int ns = g.ref_chunks;
    const size_t rows2 = (size_t)intra_dir_rows(g);

    // ---- end cell ----------------------------------------------------------------------
    int i, j;
    if (packed && NW) {
        // arg-max of matrix row `rows` (first strictly greater column, column 0 = rows*gap_ref first): the fill kernel left
        // one key per 16-column strip and lane, (H << 16) | (0xFFFF - column)
        int key = (int)(((uint32_t)(rows * gap_ref) << 16) | 0xFFFFu);
        const int nstrips = (cols + INTRA_TW - 1) / INTRA_TW;
        const uint32_t *hk = b.hrow + (size_t)duo * 2 + (slot & 1);
        for (int st = lane; st < nstrips; st += 32) key = max(key, (int)hk[(size_t)st * g.duos * 2]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(FULL, key, o));
        const int best = key >> 16, idx = 0xFFFF - (key & 0xFFFF);
        i = rows - 1;
        // pad columns (never filled) take part in the reference's arg-max: see traceback_kernel
        const int pad_cols = g.ref_length - cols, reach = min(pad_cols, rows);
        int col_max = (reach == rows && rows > 0) ? 0 : INT_MIN;  // matrix row 0
        if (reach > 0) {
            const uint32_t *bl = b.fboundary + (size_t)duo * g.rows_alloc;  // last true column, true H
            for (int r = rows - 2 - lane; r >= rows - 1 - reach && r >= 0; r -= 32) col_max = max(col_max, (int)(int16_t)(bl[r] >> lane_shift));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) col_max = max(col_max, __shfl_xor_sync(FULL, col_max, o));
        }
        j = (pad_cols > 0 && col_max > best) ? (int)meta.max_ref_pos : min((int)meta.max_ref_pos, idx);
        if (lane == 0) {
            b.scores[pair] = (int16_t)best;
            b.end_cell[2 * pair] = (int16_t)i;
            b.end_cell[2 * pair + 1] = (int16_t)j;
        }
    } else {
        i = b.end_cell[2 * pair];
        j = b.end_cell[2 * pair + 1];
    }
    const int end_i = i, end_j = j;
    const int score = (int)b.scores[pair];

    // ---- walk (every lane carries the same state) ----------------------------------------
    int final_i = end_i, final_j = end_j;
    auto walk = [&](auto &&sink) -> int {
        int i = end_i, j = end_j;
        int n_moves = 0;
        if (packed) {
            const bool in_range = i >= 0 && i < rows && j < cols;
            if (in_range && j >= 0 && (NW || score > 0)) {
                const uint2 *base = reinterpret_cast<const uint2 *>(b.fdirs) + (size_t)duo * ns * rows2;
                const uint32_t pol = (uint32_t)g.policy & 1u;
                int strip = j >> 4, k = j & 15;
                int ws = -2, wr0 = 0;  // window: rows [wr0, wr0 + 31] of strips ws and ws - 1
                uint2 wa = make_uint2(0u, 0u), wb = make_uint2(0u, 0u);
                int hval = NW ? 1 : score;
                ByteWindow wread(seq_ptr(b.raw_reads, b.read_off, pair, g.read_length), seq_end(b.raw_reads, b.read_off, g.n, g.read_length));
                ByteWindow wref(seq_ptr(b.raw_refs, b.ref_off, pair, g.ref_length), seq_end(b.raw_refs, b.ref_off, g.n, g.ref_length));
                while (true) {
                    if ((strip != ws && strip != ws - 1) || i < wr0) {
                        ws = strip;
                        wr0 = i - 31;
                        const int row = wr0 + lane;
                        if (row >= 0) {
                            wa = __ldg(base + (size_t)ws * rows2 + row);
                            if (ws > 0) wb = __ldg(base + (size_t)(ws - 1) * rows2 + row);
                        }
                    }
                    const uint2 mine = strip == ws ? wa : wb;
                    const int src = i - wr0;
                    const uint32_t wx = __shfl_sync(FULL, mine.x, src), wy = __shfl_sync(FULL, mine.y, src);
                    const int bit = lane_shift + k;
                    const uint32_t dbit = (wx >> bit) & 1u, ubit = (wy >> bit) & 1u;
                    const int code = dbit ? DIR_DIAG : ((ubit ^ pol) ? DIR_UP : DIR_LEFT);
                    sink(code, n_moves);
                    ++n_moves;
                    if (!NW) {  // SW: a zero cell is START (DefaultKernel.cpp:240-241); the value is known along the path
                        if (code == DIR_UP) hval -= sc.gap_ref;
                        else if (code == DIR_LEFT) hval -= sc.gap_read;
                        else {
                            const unsigned ca = wread.get(i) & 0xDFu, cb = wref.get(j) & 0xDFu;
                            const bool va = ca == 'A' || ca == 'C' || ca == 'G' || ca == 'T', vb = cb == 'A' || cb == 'C' || cb == 'G' || cb == 'T';
                            hval -= (va && vb) ? (ca == cb ? sc.match : sc.mismatch) : 0;
                        }
                    }
                    if (code != DIR_LEFT) --i;
                    if (code != DIR_UP) {
                        --j;
                        if (k == 0) {
                            --strip;
                            k = 15;
                        } else {
                            --k;
                        }
                    }
                    if ((i | j) < 0 || (!NW && hval <= 0)) break;
                }
            }
            // matrix column 0 (DefaultKernel.cpp:304): NW walks up to row 0, SW stops
            if (NW && i >= 0 && i < rows && j < 0) {
                for (; i >= 0; --i) {
                    sink(DIR_UP, n_moves);
                    ++n_moves;
                }
            }
        } else {
            while (true) {
                int code;
                if (i < 0 || i >= rows || j >= cols) code = DIR_START;
                else if (j < 0) code = NW ? DIR_UP : DIR_START;
                else code = (b.dirs[((size_t)(j >> 3) * g.rows_alloc + i) * g.slots + slot] >> (2 * (j & 7))) & 3;
                if (code == DIR_START) break;
                sink(code, n_moves);
                ++n_moves;
                if (code != DIR_LEFT) --i;
                if (code != DIR_UP) --j;
            }
        }
        final_i = i;
        final_j = j;
        return n_moves;
    };
    // 2-bit queue sink: 16 moves per word, written by lane 0
    uint32_t acc = 0;
    uint32_t *qbase = gq + slot;
    size_t qstride = (size_t)g.slots;
    uint32_t *qcur = qbase;
    auto queue_sink = [&](int code, int t) {
        acc |= (uint32_t)code << (2 * (t & 15));
        if ((t & 15) == 15) {
            if (lane == 0) *qcur = acc;
            qcur += qstride;
            acc = 0;
        }
    };
    int n_moves;
    if (moves_only) {
        // see traceback_kernel: CIGAR runs in walk order, or the raw queue when there are more runs than slots
        uint32_t *out = b.moves_out + (size_t)pair * (queue_words + 1);
        int cur = -1, len = 0, nruns = 0;
        auto rle_sink = [&](int code, int) {
            const int op = code == DIR_DIAG ? 0 : code == DIR_UP ? 1 : 2;
            if (op == cur) {
                ++len;
            } else {
                if (len) {
                    if (nruns < queue_words && lane == 0) out[1 + nruns] = ((uint32_t)len << 4) | (uint32_t)cur;
                    ++nruns;
                }
                cur = op;
                len = 1;
            }
        };
        n_moves = walk(rle_sink);
        if (len) {
            if (nruns < queue_words && lane == 0) out[1 + nruns] = ((uint32_t)len << 4) | (uint32_t)cur;
            ++nruns;
        }
        if (nruns <= queue_words) {
            if (lane == 0) out[0] = (uint32_t)nruns;
        } else {
            qbase = qcur = out + 1;
            qstride = 1;
            n_moves = walk(queue_sink);
            if ((n_moves & 15) && lane == 0) *qcur = acc;
            if (lane == 0) out[0] = 0x80000000u | (uint32_t)n_moves;
        }
        if (lane == 0) {
            if (b.coords) {
                int32_t *co = b.coords + 4 * (size_t)pair;
                co[0] = final_i + 1;
                co[1] = end_i + 1;
                co[2] = final_j + 1;
                co[3] = end_j + 1;
            }
            if (b.run_count) b.run_count[pair] = (uint32_t)nruns;
            b.start[pair] = (int16_t)(L - 1 - n_moves);
        }
        return;
    }
    n_moves = walk(queue_sink);
    if ((n_moves & 15) && lane == 0) *qcur = acc;
    __syncwarp();
    __threadfence_block();

    // ---- emit (see traceback_kernel) -------------------------------------------------------
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint8_t *rd = seq_ptr(b.raw_reads, b.read_off, pair, g.read_length) + end_i;
    const uint8_t *rf = seq_ptr(b.raw_refs, b.ref_off, pair, g.ref_length) + end_j;
    uint8_t *oa, *ob;
    if (compact) {
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(b.compact_cursor, 2ull * (unsigned long long)(n_moves + 1));
            b.compact_off[pair] = (uint32_t)base;
        }
        base = __shfl_sync(FULL, base, 0);
        oa = b.aln_compact + (size_t)base + n_moves;
        ob = oa + n_moves + 1;
    } else {
        oa = b.aln_read + (size_t)pair * L + (L - 1);
        ob = b.aln_ref + (size_t)pair * L + (L - 1);
    }
    for (int t0 = 0; t0 < n_moves; t0 += 32) {
        const int t = t0 + lane;
        const bool valid = t < n_moves;
        const uint32_t word = valid ? qbase[(size_t)(t >> 4) * qstride] : 0u;
        const int code = (word >> (2 * (t & 15))) & 3;
        const bool takes_r = valid && code != DIR_LEFT, takes_f = valid && code != DIR_UP;
        const unsigned mr = __ballot_sync(FULL, takes_r), mf = __ballot_sync(FULL, takes_f);
        uint8_t a = '-', c = '-';
        if (takes_r) a = *(rd - __popc(mr & lt_mask));
        if (takes_f) c = *(rf - __popc(mf & lt_mask));
        if (valid && (compact || t <= L - 2)) {
            oa[-1 - t] = a;
            ob[-1 - t] = c;
        }
        rd -= __popc(mr);
        rf -= __popc(mf);
    }
    if (lane == 0) {
        b.start[pair] = (int16_t)(L - 1 - n_moves);
        if (compact || L >= 1) {
            *oa = 0;
            *ob = 0;
        }
    }
}

// Packed entry points: pair i's runs leave the traceback in walk order in its fixed slot (or as the raw 2-bit
// queue when there were more runs than slots); here they become the forward-order BAM CIGAR at run_offs[i].
__global__ void __launch_bounds__(256) cigar_compact_kernel(int n, int queue_words, const uint32_t *__restrict__ moves,
                                                            const uint32_t *__restrict__ run_offs, uint32_t *__restrict__ out,
                                                            size_t out_cap) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n) return;
    // a chunk whose runs do not fit the block is replayed on the host instead (the caller checks run_offs[n])
    if ((size_t)run_offs[pair + 1] > out_cap) return;
    const uint32_t *q = moves + (size_t)pair * (queue_words + 1);
    uint32_t *o = out + run_offs[pair];
    const uint32_t head = q[0];
    ++q;
    if (!(head & 0x80000000u)) {
        for (int r = (int)head - 1; r >= 0; --r) *o++ = q[r];
        return;
    }
    const int n_moves = (int)(head & 0x7FFFFFFFu);
    int cur = -1, len = 0;
    for (int t = n_moves - 1; t >= 0; --t) {
        const int code = (q[t >> 4] >> (2 * (t & 15))) & 3;
        const int op = code == DIR_DIAG ? 0 : code == DIR_UP ? 1 : 2;
        if (op == cur) {
            ++len;
        } else {
            if (len) *o++ = ((uint32_t)len << 4) | (uint32_t)cur;
            cur = op;
            len = 1;
        }
    }
    if (len) *o++ = ((uint32_t)len << 4) | (uint32_t)cur;
}

}  // namespace

size_t cigar_compact_scratch_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, n + 1);
    return bytes + 256;
}

int launch_cigar_compact(int n, int queue_words, const uint32_t *moves, const uint32_t *run_count, uint32_t *run_offs,
                         uint32_t *cigar_out, size_t out_cap_words, void *scratch, size_t scratch_bytes, cudaStream_t stream) {
    if (n <= 0) return 0;
    cub::DeviceScan::ExclusiveSum(scratch, scratch_bytes, run_count, run_offs, n + 1, stream);
    cigar_compact_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, queue_words, moves, run_offs, cigar_out, out_cap_words);
    return 1;  // our kernel; the scan is the library's
}

size_t traceback_queue_words(int read_length, int ref_length) { return (size_t)(read_length + ref_length + 15) / 16 + 1; }

// The per-thread move queues live in dynamic shared memory while they fit beside the kernel's static arrays
// (48 KB per block without opt-in); beyond that the caller provides slots * queue_words words of global memory.
bool traceback_needs_global_queue(int read_length, int ref_length) {
    return traceback_queue_words(read_length, ref_length) * TB_THREADS * sizeof(uint32_t) + 8 * TB_THREADS * sizeof(int) > 48 * 1024;
}

// long shapes may run the warp-per-pair traceback (chunks filled by the intra-task kernels), whose queues are always global
bool traceback_wants_global_queue(int read_length, int ref_length) {
    return traceback_needs_global_queue(read_length, ref_length) || (read_length >= 256 && ref_length >= 1024);
}

int launch_traceback(const ChunkGeom &g, const ChunkBuffers &b, int mode, const Scoring &sc, uint32_t *global_queue,
                     cudaStream_t stream) {
    if (g.n <= 0) return 0;
    const int qw = (int)traceback_queue_words(g.read_length, g.ref_length);
    if (g.intra) {  // long pairs: one warp per pair, queues in global memory
        const int blocks = (int)(((long long)g.n * 32 + TB_THREADS - 1) / TB_THREADS);
        if (mode == MODE_NW_ALIGN) traceback_long_kernel<true><<<blocks, TB_THREADS, 0, stream>>>(g, b, sc, global_queue, qw);
        else traceback_long_kernel<false><<<blocks, TB_THREADS, 0, stream>>>(g, b, sc, global_queue, qw);
        return 1;
    }
    const int blocks = (g.n + TB_THREADS - 1) / TB_THREADS;
    const size_t smem = (size_t)qw * TB_THREADS * sizeof(uint32_t);
    const bool use_shared = !traceback_needs_global_queue(g.read_length, g.ref_length) && !b.moves_out;
    uint32_t *gq = use_shared ? nullptr : global_queue;
    if (mode == MODE_NW_ALIGN) traceback_kernel<true><<<blocks, TB_THREADS, use_shared ? smem : 0, stream>>>(g, b, sc, gq, qw);
    else traceback_kernel<false><<<blocks, TB_THREADS, use_shared ? smem : 0, stream>>>(g, b, sc, gq, qw);
    return 1;
}

}  // namespace va
