// va_internal.h -- host-visible declarations shared by the kernel launchers (va_kernels.cu)
// and the C ABI / staging layer (va_cabi.cu).  Nothing here is exported.
#ifndef VA_INTERNAL_H
#define VA_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

namespace va {

// Base codes written by the prep kernel.  0..3 are the four bases that can score;
// N and "anything else" (including the '\0' pad and bytes >= 0x80) always score 0 but are
// told apart because the Default/OpenCL kernels count N as a valid character when they
// look for the end of a sequence (DefaultKernel.cpp:308,348) while SSE/AVX do not
// (SSEKernel.cpp:514-518,673-677).
enum : int { CODE_A = 0, CODE_C = 1, CODE_G = 2, CODE_T = 3, CODE_N = 4, CODE_OTHER = 5 };
// Row index only (never a base code): a row in front of a lane's first matrix row.  A duo whose reads differ
// in length is END-aligned in packed NW align -- the shorter lane starts max(rows) - rows sweep rows late --
// and in the shifted recurrence a row whose table is all zero just hands matrix row 0 down (va_nw.cu).
enum : int { CODE_PRE = 6 };

// direction codes, shared with the traceback kernel (same numbering as SSEKernel.h:28-31)
enum : int { DIR_START = 0, DIR_UP = 1, DIR_LEFT = 2, DIR_DIAG = 3 };

enum : int { MODE_SW_SCORE = 0, MODE_NW_SCORE = 1, MODE_SW_ALIGN = 2, MODE_NW_ALIGN = 3 };

struct Scoring {
    int match, mismatch, gap_read, gap_ref;
};

// Per-pair DP extents and NW end-of-sequence markers, produced by the prep kernel.
struct __align__(16) PairMeta {
    int16_t rows;          // DP rows to fill   (<= read_length)
    int16_t cols;          // DP columns to fill (<= ref_length)
    int16_t max_read_pos;  // NW align: index of the first invalid read char - 1 (policy dependent)
    int16_t max_ref_pos;   // NW align: same for the ref
    int16_t true_rows;     // 1 + index of the last ACGT base of the read
    int16_t true_cols;     // same for the ref
    int16_t flags;         // bit 0: a non-ACGT byte inside [0,true_rows) of the read; bit 1: same for ref
    int16_t pad;
};

// Geometry of one chunk's device buffers.  "slot" = position of a pair inside the chunk.
struct ChunkGeom {
    int n;             // pairs in the chunk
    int slots;         // n rounded up (multiple of 64): stride of every slot-interleaved array
    int read_length;   // batch-wide padded lengths (bytes per raw sequence)
    int ref_length;
    int read_chunks;   // ceil(read_length/16): uint4 code chunks per read
    int ref_chunks;
    int rows_alloc;    // rows of the direction matrix / boundary column that are allocated
    int segs;          // ceil(ref_length/8): direction half-words per row
    int fast_tw;       // column-strip width of the packed 16-bit kernels for this call, 0 = not eligible
    int duos;          // slots / 2: stride of the arrays the packed kernels index by pair-of-pairs
    int solo;          // 1: the packed inter-task kernels also take single slots whose duo is not fast (va_fast.cuh)
    int policy;        // traceback pointer policy of the call (0 Default/OpenCL, 1 SSE/AVX); align modes only
    int affine;        // 1: affine-gap variant (general kernel only): a gap of length L costs gap_open + L * gap_{read,ref}
    int gap_open;      // <= 0
    int intra;         // 1: the chunk's packed duos are computed by the intra-task kernels (va_intra.cu): 16-column strips,
                       // directions at [duo][strip][row], boundary at [duo][row], no end-aligned NW duos
    int inband;        // 1 (NW align, inter-task packed kernels): values are carried as 4V + tag and the direction words hold
                       // 2-bit tags instead of two bit planes (va_nw.cu); the boundary column holds 4V
};

// Constants of the packed (two pairs per thread, s16x2) kernels, built on the host per call.
struct FastConsts {
    uint32_t tab[8];   // tab[read_code]: four 8-bit substitution scores vs ref A,C,G,T (NW align: minus gap_ref)
    uint32_t gF2, gR2; // gap scores replicated in both 16-bit lanes
    uint32_t dFR2;     // SW: boundary ("left + gR") -> diagonal conversion: gap_ref - gap_read (align), -gap_read (score)
    int gF, gR;
    // SW align only (see va_fast.cu): biased initial values and key constants
    uint32_t swa_l0, swa_g0, swa_k32;
    uint32_t swa_zero;  // the biased value of a zero cell (SSE/AVX policy: "not START" is h >= 0)
    int swa_off;
};

struct ChunkBuffers {
    const uint8_t *raw_reads;  // [n][read_length], or -- with read_off -- the chunk's sequences back to back
    const uint8_t *raw_refs;   // [n][ref_length]
    // offset-addressed input (packed entry points): sequence i = raw_reads + (read_off[i] - read_off[0]), length
    // read_off[i+1] - read_off[i] <= read_length; n+1 entries, the caller's own offsets (any base).  NULL = fixed stride.
    const int64_t *read_off;
    const int64_t *ref_off;
    uint4 *code_reads;         // [read_chunks][slots]
    uint4 *code_refs;          // [ref_chunks][slots]
    uint4 *row_idx;            // [read_chunks][duos]: per sweep row 7*code(slot 2u) + code(slot 2u+1), 16 rows per word
                               // (the packed NW kernels' index into their table of score-table pairs)
    PairMeta *meta;            // [slots]
    int32_t *pair_of;          // [slots] pair (position in the caller's batch) computed in this slot, -1 = padding
    int32_t *boundary;         // general kernel: [rows_alloc][slots] right edge of the previous column strip
    uint32_t *fboundary;       // packed kernels: [rows_alloc][duos] (inter-task) or [duos][rows_alloc] (intra-task);
                               // after a packed NW-align fill it holds the last true column (read by the traceback)
    uint16_t *dirs;            // general kernel: [segs][rows_alloc][slots] half-words, 2 bits per cell, 8 cells each
    uint4 *fdirs;              // packed kernel:  [strip][row pair][group][duos], see va_fast.cuh
                               // (separate regions: one chunk can hold pairs of both kinds)
    uint2 *fdirs_z;            // packed SW align under the SSE/AVX policy: third plane "cell is not START" (h >= 0), same
                               // indexing as fdirs, .x = even row / .y = odd row of the pair, low 16 bits lane A; else NULL
    int32_t *solo_list;        // [slots] slots the packed kernels take on their own (va_fast.cuh), written by the prep
    int32_t *solo_count;       // kernel in no particular order; *solo_count entries
    int32_t *boundary_e;       // affine variant: [rows_alloc][slots] E (gap-in-read state) of the previous strip's right edge
    uint32_t *dirs4;           // affine variant: [segs][rows_alloc][slots] words, 4 bits per cell (H source, E opened, F opened), 8 cells each
    uint32_t *hrow;            // packed NW align: [strip][duo][2] arg-max key of the last valid matrix row per strip and lane
    int16_t *scores;           // [n]
    int16_t *end_cell;         // [n][2]
    // traceback outputs
    uint8_t *aln_read;         // [n][read_length+ref_length]
    uint8_t *aln_ref;
    int16_t *start;            // [n]
    uint32_t *moves_out;       // when set: the traceback stops after the walk and leaves pair i's moves at
                               // moves_out[i * (queue_words + 1) ..): CIGAR runs in walk order or the raw 2-bit queue
                               // (va_traceback.cu); start[i] = L - 1 - moves as usual, aln_read / aln_ref are not written
    // packed entry points, beside moves_out (all optional)
    int32_t *coords;           // [n][4] aligned region in sequence coordinates: read_begin, read_end, ref_begin, ref_end
    uint32_t *run_count;       // [n] CIGAR runs of the pair (also of pairs whose moves left as the raw queue)
    // compact string output (host pipeline of the legacy boundary): when aln_compact is set the two strings of pair i,
    // each NUL terminated, (moves + 1) bytes, lie back to back at aln_compact + compact_off[i]; *compact_cursor
    // (zeroed by the caller) ends up as the bytes used.  aln_read / aln_ref are not written then.
    uint8_t *aln_compact;
    uint32_t *compact_off;
    unsigned long long *compact_cursor;
    unsigned long long *cell_count;  // device counter: DP cells computed
};

// Launchers (va_kernels.cu).  All asynchronous on `stream`; return the number of kernels launched.
// staging (va_prep.cu)
size_t prep_scratch_bytes(int slots, int read_length, int ref_length);
int launch_prep(const ChunkGeom &g, const ChunkBuffers &b, int mode, int policy, const Scoring &sc, void *scratch,
                size_t scratch_bytes, cudaStream_t stream);
int launch_fill_general(const ChunkGeom &g, const ChunkBuffers &b, int mode, int policy, const Scoring &sc,
                        cudaStream_t stream);
// packed kernels (va_fast.cu)
// intra: the intra-task kernels' domain (SW align without the 16-bit best-cell key, NW modes un-shifted)
bool fast_scoring_ok(int mode, int policy, const Scoring &sc, int read_length, int ref_length, bool intra = false);
int fast_pick_tw(int mode, int ref_length, int policy = 0);
size_t fast_dirs_bytes_per_row_per_slot(int ref_length);
FastConsts make_fast_consts(int mode, const Scoring &sc, bool inband = false);
// NW align on the packed inter-task kernels: may the values be carried as 4V + tag (va_nw.cu)?
bool fast_inband_ok(int mode, int policy, const Scoring &sc, int read_length, int ref_length);
int launch_fill_fast(const ChunkGeom &g, const ChunkBuffers &b, int mode, const Scoring &sc, cudaStream_t stream);
// packed kernels of the NW modes, shifted recurrence (va_nw.cu); reached through launch_fill_fast
int launch_fill_nw(const ChunkGeom &g, const ChunkBuffers &b, int mode, const FastConsts &fc, cudaStream_t stream);
// intra-task (warp per pair-of-pairs) kernel for few, long pairs (va_intra.cu)
bool intra_preferred(int mode, int n_pairs, int read_length, int ref_length, int sm_count);
int launch_fill_intra(const ChunkGeom &g, const ChunkBuffers &b, int mode, const FastConsts &fc, cudaStream_t stream);
size_t traceback_queue_words(int read_length, int ref_length);
bool traceback_needs_global_queue(int read_length, int ref_length);
bool traceback_wants_global_queue(int read_length, int ref_length);
int launch_traceback(const ChunkGeom &g, const ChunkBuffers &b, int mode, const Scoring &sc, uint32_t *global_queue,
                     cudaStream_t stream);
// packed entry points: per-pair CIGAR runs (walk order, fixed slots) -> one forward-order block (va_traceback.cu).
// run_count has n+1 entries (the last one 0); run_offs receives its exclusive prefix sum (run_offs[n] = total runs).
size_t cigar_compact_scratch_bytes(int n);
int launch_cigar_compact(int n, int queue_words, const uint32_t *moves, const uint32_t *run_count, uint32_t *run_offs,
                         uint32_t *cigar_out, size_t out_cap_words, void *scratch, size_t scratch_bytes, cudaStream_t stream);
int launch_int_peak(int kind, int sm_count, int iters, unsigned int *sink, cudaStream_t stream, double *lane_ops);

}  // namespace va
#endif
