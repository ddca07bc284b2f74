// va_traceback.cu -- per-pair traceback over the 2-bit direction matrix and output of the two
// gapped strings, right-aligned in read_length+ref_length byte blocks like the reference's
// calc_alignment_* (DefaultKernel.cpp:391-525; SSEKernel.cpp:729-1005;
// alignment_kernels.cl:146-192,370-414).
//
// One thread per pair, two phases, so that the only dependent chain through HBM is the walk:
//   walk  follow the pointers from the end cell to START, one direction word per step (L2/HBM
//         latency bound, hidden by occupancy); the moves go into a per-thread 2-bit queue in
//         shared memory (global memory when the sequences are too long for that);
//   emit  replay the queue: fetch the read/ref bytes it names and write both strings backwards,
//         four characters per 32-bit store, so every output sector is written whole.
// The packed fill kernel leaves NW's end-cell decision (arg-max of the last valid row,
// DefaultKernel.cpp:352-355,381-387) to this kernel: `hrow` holds that row.
#include <climits>

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

// Backwards byte STREAM over one raw sequence for the emit phase, which consumes each sequence strictly
// from its end cell down to index 0: aligned 16-byte window loads, the window's words shifted down as
// they are used up, one shift per byte -- no address arithmetic and no selects per byte.  Windows are
// loaded lazily (only when a byte of them is needed), so nothing below the buffer is touched; the first
// window is assembled byte-wise when it would reach past `limit` (end of the whole raw buffer).
struct BackStream {
    const uint4 *p;  // next (lower) window
    uint4 w;
    uint32_t cur;
    int nbytes, nwords;
    __device__ __forceinline__ void pop_word() {
        cur = w.w;
        w.w = w.z;
        w.z = w.y;
        w.y = w.x;
        --nwords;
    }
    // the first byte returned is base[idx]; idx < 0 makes an empty stream (never read)
    __device__ __forceinline__ void init(const uint8_t *base, int idx, const uint8_t *limit) {
        nbytes = nwords = 0;
        cur = 0;
        w = make_uint4(0, 0, 0, 0);
        p = nullptr;
        if (idx < 0) return;
        const uintptr_t a = reinterpret_cast<uintptr_t>(base + idx), a16 = a & ~(uintptr_t)15;
        if (a16 + 16 <= reinterpret_cast<uintptr_t>(limit)) {
            w = *reinterpret_cast<const uint4 *>(a16);
        } else {  // last window of the buffer: only the bytes that exist
            uint32_t t[4] = {0, 0, 0, 0};
            for (uintptr_t q = a16; q <= a; ++q) t[(q - a16) >> 2] |= (uint32_t)*reinterpret_cast<const uint8_t *>(q) << (8 * ((q - a16) & 3));
            w = make_uint4(t[0], t[1], t[2], t[3]);
        }
        p = reinterpret_cast<const uint4 *>(a16) - 1;
        nwords = 4;
        const int o = (int)(a & 15);
        for (int d = 3; d > (o >> 2); --d) pop_word();  // words above the start are not part of the stream
        pop_word();
        cur <<= 8 * (3 - (o & 3));
        nbytes = (o & 3) + 1;
    }
    __device__ __forceinline__ uint32_t next() {
        if (nbytes == 0) {
            if (nwords == 0) {
                w = *p;
                --p;
                nwords = 4;
            }
            pop_word();
            nbytes = 4;
        }
        const uint32_t b = cur >> 24;
        cur <<= 8;
        --nbytes;
        return b;
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
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= g.n) return;
    const bool moves_only = b.moves_out != nullptr;
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
    const int L = g.read_length + g.ref_length;
    const PairMeta meta = b.meta[slot];
    const int rows = meta.rows, cols = meta.cols;
    const int mode = NW ? MODE_NW_ALIGN : MODE_SW_ALIGN;
    const bool packed = slot_owner(g, mode, slot, b.meta[slot & ~1], b.meta[slot | 1]) != OWN_NONE;
    const int duo = slot >> 1, lane_shift = (slot & 1) * 16;
    const int pair = b.pair_of[slot];  // raw bytes and results are indexed in the caller's pair order

    // ---- end cell ----------------------------------------------------------------------
    int i, j;
    if (packed && NW) {
        // first strictly greater column of matrix row `rows`, starting from column 0 = rows*gap_ref.
        // hrow holds that row in the fill kernel's shifted form V = H - gap_ref*I - gap_read*J (va_nw.cu)
        int best = rows * gap_ref, idx = 0;
        const uint32_t *hr = b.hrow + duo;
        for (int c = 0; c < cols; ++c) {
            const int h = (int)(int16_t)(hr[(size_t)c * g.duos] >> lane_shift) + rows * gap_ref + (c + 1) * sc.gap_read;
            if (h > best) {
                best = h;
                idx = c;
            }
        }
        i = rows - 1;
        // Pad columns (past `cols`, never filled) take part in the arg-max of the reference.  With both
        // gaps <= 0 one of them beats the best true cell exactly when a value of the last true column in
        // one of the min(pad columns, rows) matrix rows above comes down a zero-score diagonal; the
        // arg-max is then past max_ref_pos and gets clipped to it (DefaultKernel.cpp:387).
        const int pad_cols = g.ref_length - cols, reach = min(pad_cols, rows);
        int col_max = (reach == rows && rows > 0) ? 0 : INT_MIN;  // matrix row 0
        if (reach > 0) {
            const uint32_t *bl = b.fboundary + duo;  // last true column (matrix column `cols`), shifted
            for (int r = rows - 2; r >= rows - 1 - reach && r >= 0; --r)
                col_max = max(col_max, (int)(int16_t)(bl[(size_t)r * g.duos] >> lane_shift) + (r + 1) * gap_ref + cols * sc.gap_read);
        }
        j = (pad_cols > 0 && col_max > best) ? (int)meta.max_ref_pos : min((int)meta.max_ref_pos, idx);
        b.end_cell[2 * pair] = (int16_t)i;
        b.end_cell[2 * pair + 1] = (int16_t)j;
        b.scores[pair] = (int16_t)best;
    } else {
        i = b.end_cell[2 * pair];
        j = b.end_cell[2 * pair + 1];
    }
    const int end_i = i, end_j = j;

    // ---- walk ----------------------------------------------------------------------------
    // One walk from the end cell to START; every move goes to `sink(code, t)` (t = moves so far).
    auto walk = [&](auto &&sink) -> int {
        int i = end_i, j = end_j;
        int n_moves = 0;
        if (packed) {
            const int tw = g.fast_tw, ng = fast_groups(tw);
            int strip = j >= 0 ? j / tw : 0, k = j >= 0 ? j - strip * tw : 0;
            // NW: a partial last strip keeps its true columns in the LAST registers of the strip (va_nw.cu)
            int kmin = 0;
            if (NW && cols > 0 && strip == (cols - 1) / tw) {
                kmin = tw - (cols - strip * tw);
                k += kmin;
            }
            const size_t pair_step = (size_t)ng * g.duos, strip_step = (size_t)fast_row_pairs(g) * pair_step;
            const uint4 *p = b.fdirs + (size_t)strip * strip_step + (size_t)(max(i, 0) >> 1) * pair_step + duo;
            int have_pair = -1, have_strip = -1, have_grp = -1;
            uint4 w = make_uint4(0, 0, 0, 0);
            // SW: the packed fill stores the pointer a cell would have in NW; a cell whose value is 0 is
            // START (DefaultKernel.cpp:240-241).  The value is known along the path: it starts at the best
            // score and every move gives back what it added.
            int hval = NW ? 1 : (int)b.scores[pair];
            ByteWindow wread(b.raw_reads + (size_t)pair * g.read_length, b.raw_reads + (size_t)g.n * g.read_length);
            ByteWindow wref(b.raw_refs + (size_t)pair * g.ref_length, b.raw_refs + (size_t)g.n * g.ref_length);
            while (true) {
                int code;
                if (i < 0 || i >= rows || j >= cols) code = DIR_START;
                else if (j < 0) code = NW ? DIR_UP : DIR_START;  // matrix column 0 (DefaultKernel.cpp:304)
                else if (!NW && hval <= 0) code = DIR_START;
                else {
                    const int grp = k >> 4;
                    if ((i >> 1) != have_pair || strip != have_strip || grp != have_grp) {
                        w = p[(size_t)grp * g.duos];  // two rows x 16 columns x both lanes
                        have_pair = i >> 1;
                        have_strip = strip;
                        have_grp = grp;
                    }
                    const int bit = lane_shift + (k & 15);
                    const uint32_t diag_plane = (i & 1) ? w.z : w.x, up_plane = (i & 1) ? w.w : w.y;
                    code = ((diag_plane >> bit) & 1) ? DIR_DIAG : (((up_plane >> bit) & 1) ? DIR_UP : DIR_LEFT);
                }
                if (code == DIR_START) break;
                sink(code, n_moves);
                ++n_moves;
                if (!NW) {
                    if (code == DIR_UP) hval -= sc.gap_ref;
                    else if (code == DIR_LEFT) hval -= sc.gap_read;
                    else {
                        const unsigned ca = wread.get(i) & 0xDFu, cb = wref.get(j) & 0xDFu;
                        const bool va = ca == 'A' || ca == 'C' || ca == 'G' || ca == 'T', vb = cb == 'A' || cb == 'C' || cb == 'G' || cb == 'T';
                        hval -= (va && vb) ? (ca == cb ? sc.match : sc.mismatch) : 0;
                    }
                }
                if (code != DIR_LEFT) {
                    if ((i & 1) == 0) p -= pair_step;  // leaving an even row: the row above is in the previous word
                    --i;
                }
                if (code != DIR_UP) {
                    --j;
                    if (--k < kmin) {
                        k = tw - 1;
                        kmin = 0;
                        --strip;
                        p -= strip_step;
                    }
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
        return n_moves;
    };
    // 2-bit queue sink: 16 moves per word
    uint32_t acc = 0;
    auto queue_sink = [&](int code, int t) {
        acc |= (uint32_t)code << (2 * (t & 15));
        if ((t & 15) == 15) {
            mq.q[(size_t)(t >> 4) * mq.stride] = acc;
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
            mq.q = out + 1;
            mq.stride = 1;
            n_moves = walk(queue_sink);
            if (n_moves & 15) mq.q[(size_t)(n_moves >> 4)] = acc;
            out[0] = 0x80000000u | (uint32_t)n_moves;
        }
    } else {
        n_moves = walk(queue_sink);
        if (n_moves & 15) mq.q[(size_t)(n_moves >> 4) * mq.stride] = acc;
    }

    if (moves_only) {
        b.start[pair] = (int16_t)(g.read_length + g.ref_length - 1 - n_moves);
        return;
    }
    // ---- emit ----------------------------------------------------------------------------
    BackStream read, ref;
    read.init(b.raw_reads + (size_t)pair * g.read_length, end_i, b.raw_reads + (size_t)g.n * g.read_length);
    ref.init(b.raw_refs + (size_t)pair * g.ref_length, end_j, b.raw_refs + (size_t)g.n * g.ref_length);
    uint8_t *oa = b.aln_read + (size_t)pair * L;
    uint8_t *ob = b.aln_ref + (size_t)pair * L;
    const int start = L - 1 - n_moves;  // may be negative only when every move was a gap (never with gap scores < 0)
    b.start[pair] = (int16_t)start;
    if (L >= 1) {
        oa[L - 1] = 0;
        ob[L - 1] = 0;
    }
    int pos = L - 2;
    int t = 0;  // moves replayed so far
    uint32_t mw = n_moves ? mq.q[0] : 0;
    int mleft = 16;  // moves left in mw
    size_t mnext = mq.stride;
    auto next_move = [&]() -> int {
        if (mleft == 0) {
            mw = mq.q[mnext];
            mnext += mq.stride;
            mleft = 16;
        }
        const int code = mw & 3;
        mw >>= 2;
        --mleft;
        ++t;
        return code;
    };
    auto one_byte = [&]() {
        const int code = next_move();
        uint8_t a = '-', c = '-';
        if (code != DIR_LEFT) a = (uint8_t)read.next();
        if (code != DIR_UP) c = (uint8_t)ref.next();
        if (pos >= 0) {
            oa[pos] = a;
            ob[pos] = c;
        }
        --pos;
    };
    // Both output rows start at pair*L from (at least 256-byte aligned) buffers, so they share their
    // alignment: bytes down to a 4-byte boundary, words down to a 16-byte boundary, then 16 moves
    // per pair of 16-byte stores, then the remainder the same way back down.
    const uintptr_t base_a = reinterpret_cast<uintptr_t>(oa), base_b = reinterpret_cast<uintptr_t>(ob);
    auto four_moves = [&](uint32_t &wa, uint32_t &wb) {
        wa = 0;
        wb = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int code = next_move();
            uint32_t a = '-', c = '-';
            if (code != DIR_LEFT) a = read.next();
            if (code != DIR_UP) c = ref.next();
            wa |= a << (8 * (3 - r));
            wb |= c << (8 * (3 - r));
        }
    };
    auto one_word = [&]() {
        uint32_t wa, wb;
        four_moves(wa, wb);
        *reinterpret_cast<uint32_t *>(oa + pos - 3) = wa;
        *reinterpret_cast<uint32_t *>(ob + pos - 3) = wb;
        pos -= 4;
    };
    if ((base_a & 15) == (base_b & 15)) {
        while (t < n_moves && pos >= 0 && ((base_a + pos) & 3) != 3) one_byte();
        while (n_moves - t >= 4 && pos >= 3 && ((base_a + pos) & 15) != 15) one_word();
        while (n_moves - t >= 16 && pos >= 15) {
            uint32_t wa[4], wb[4];
#pragma unroll
            for (int q = 3; q >= 0; --q) four_moves(wa[q], wb[q]);  // highest addresses first
            *reinterpret_cast<uint4 *>(oa + pos - 15) = make_uint4(wa[0], wa[1], wa[2], wa[3]);
            *reinterpret_cast<uint4 *>(ob + pos - 15) = make_uint4(wb[0], wb[1], wb[2], wb[3]);
            pos -= 16;
        }
        while (n_moves - t >= 4 && pos >= 3) one_word();
    }
    while (t < n_moves) one_byte();
}

}  // namespace

// Shared-memory move queue while it fits (<= 48 KB per block), else `global_queue`
// (slots * queue_words words, provided by the caller).
size_t traceback_queue_words(int read_length, int ref_length) { return (size_t)(read_length + ref_length + 15) / 16 + 1; }

int launch_traceback(const ChunkGeom &g, const ChunkBuffers &b, int mode, const Scoring &sc, uint32_t *global_queue,
                     cudaStream_t stream) {
    if (g.n <= 0) return 0;
    const int blocks = (g.n + TB_THREADS - 1) / TB_THREADS;
    const int qw = (int)traceback_queue_words(g.read_length, g.ref_length);
    const size_t smem = (size_t)qw * TB_THREADS * sizeof(uint32_t);
    const bool use_shared = smem <= 48 * 1024 && !b.moves_out;
    uint32_t *gq = use_shared ? nullptr : global_queue;
    if (mode == MODE_NW_ALIGN) traceback_kernel<true><<<blocks, TB_THREADS, use_shared ? smem : 0, stream>>>(g, b, sc, gq, qw);
    else traceback_kernel<false><<<blocks, TB_THREADS, use_shared ? smem : 0, stream>>>(g, b, sc, gq, qw);
    return 1;
}

}  // namespace va
