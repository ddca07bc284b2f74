// va_fast.cuh -- device-side pieces shared by the packed (s16x2) fill kernels and the
// traceback kernel: which pairs the packed path owns, and how its direction bits are laid out.
#ifndef VA_FAST_CUH
#define VA_FAST_CUH

#include "va_device.cuh"
#include "va_internal.h"

namespace va {

// A "duo" is two neighbouring slots (2u, 2u+1) computed by ONE thread in the two 16-bit lanes
// of every register.  The packed kernels own a duo iff this predicate holds; otherwise the
// general kernel computes both pairs.  Both kernels evaluate it, so no flag is exchanged.
//   - the per-column PRMT selector cannot express "scores 0", so no non-ACGT byte may sit inside
//     the ref columns that are filled (non-ACGT READ bytes are fine: their row table is all zero);
//   - both lanes sweep the same columns; rows may differ: for SW and the score modes trailing rows
//     that can only score 0 are neutral; NW align end-aligns the lanes (CODE_PRE rows, va_internal.h).
__device__ __forceinline__ bool duo_is_fast(const ChunkGeom &g, int mode, int slot_a, const PairMeta &a, const PairMeta &b) {
    if (g.fast_tw == 0 || slot_a + 1 >= g.n) return false;
    if ((a.flags & 2) || (b.flags & 2)) return false;
    if (a.cols != b.cols || a.cols <= 0) return false;
    if (a.cols != a.true_cols || b.cols != b.true_cols) return false;  // no padded / N tail inside the sweep
    // SSE/AVX policy, SW: DIAG only between two ACGT bases (SSEKernel.cpp:366-379) -- the planes cannot say "masked",
    // so no non-ACGT byte inside the swept read rows either (NW rows end at the first one anyway)
    // (a row past the read's last base -- the one row an empty read is given -- counts as one too)
    if (mode == MODE_SW_ALIGN && g.policy == 1 && (((a.flags | b.flags) & 1) || a.rows > a.true_rows || b.rows > b.true_rows)) return false;
    // NW align's end-cell rule reads each lane's last valid row from the registers after the sweep: lanes
    // of different read lengths are end-aligned (nw_row_offset), so both need at least one row
    // (the intra-task kernel starts both lanes at row 0 and captures each lane's own last row)
    if (mode == MODE_NW_ALIGN) return a.rows > 0 && b.rows > 0;
    return max(a.rows, b.rows) > 0;
}

// A slot whose duo is NOT fast can still be computed by the packed inter-task kernels on its own
// ("solo"): a thread sweeps that slot's extents, the other 16-bit lane computes along on whatever its
// registers hold (lanes never mix), and only the owner's halves of the shared words are stored.  That
// keeps the leftovers of the length bucketing (a run of equal extents with an odd number of pairs,
// the last slot of an odd batch, neighbours with different ref lengths) off the 32-bit general kernel.
__device__ __forceinline__ bool solo_ok(const ChunkGeom &g, int slot, const PairMeta &m) {
    if (g.fast_tw == 0 || !g.solo || slot >= g.n) return false;
    if (m.flags & 2) return false;
    if (g.policy == 1 && ((m.flags & 1) || m.rows > m.true_rows)) return false;  // (see duo_is_fast; g.policy is 0 in the score modes)
    return m.cols > 0 && m.cols == m.true_cols && m.rows > 0;
}

enum : int { OWN_NONE = 0, OWN_DUO = 1, OWN_SOLO = 2 };

// Which packed path computes `slot`; a = meta[slot & ~1], b = meta[slot | 1].  The fill kernels, the
// general kernel and the traceback kernel all evaluate this, so no flag is exchanged.
__device__ __forceinline__ int slot_owner(const ChunkGeom &g, int mode, int slot, const PairMeta &a, const PairMeta &b) {
    if (duo_is_fast(g, mode, slot & ~1, a, b)) return OWN_DUO;
    return solo_ok(g, slot, (slot & 1) ? b : a) ? OWN_SOLO : OWN_NONE;
}

// Work item of a packed inter-task kernel thread: a duo (both lanes) or one solo slot (one lane).
struct FastWork {
    int own;   // OWN_*
    int duo;   // index into the [duo]-strided arrays
    int lane;  // solo: which 16-bit lane is this thread's (0 = slot 2*duo, 1 = slot 2*duo+1)
    int rows, cols;
    PairMeta ma, mb;
};

// duo kernels: thread t takes duo t
__device__ __forceinline__ FastWork fast_work_duo(const ChunkGeom &g, const PairMeta *meta, int mode, int duo) {
    FastWork w;
    w.own = OWN_NONE;
    w.lane = 0;
    w.duo = duo;
    w.rows = w.cols = 0;
    const int slot_a = 2 * duo;
    if (slot_a + 1 >= g.n) return w;
    w.ma = meta[slot_a];
    w.mb = meta[slot_a + 1];
    if (duo_is_fast(g, mode, slot_a, w.ma, w.mb)) {
        w.own = OWN_DUO;
        w.rows = max((int)w.ma.rows, (int)w.mb.rows);
        w.cols = w.ma.cols;
    }
    return w;
}

// solo kernels: a grid-stride loop over the list the prep kernel compiled (ChunkBuffers::solo_list)
__device__ __forceinline__ FastWork fast_work_solo(const PairMeta *meta, int slot) {
    FastWork w;
    w.own = OWN_SOLO;
    w.lane = slot & 1;
    w.duo = slot >> 1;
    w.ma = meta[slot & ~1];
    w.mb = meta[slot | 1];  // slots are padded to a multiple of 64: always readable
    const PairMeta &me = w.lane ? w.mb : w.ma;
    w.rows = me.rows;
    w.cols = me.cols;
    return w;
}

// Stores into words two solo threads may share: a duo thread writes the word, a solo thread its half.
template <bool SOLO>
__device__ __forceinline__ void store_lanes(uint32_t *p, uint32_t v, const FastWork &w) {
    if (!SOLO) *p = v;
    else reinterpret_cast<uint16_t *>(p)[w.lane] = (uint16_t)(v >> (16 * w.lane));
}
template <bool SOLO>
__device__ __forceinline__ void store_lanes(uint4 *p, uint4 v, const FastWork &w) {
    if (!SOLO) {
        *p = v;
    } else {
        uint16_t *h = reinterpret_cast<uint16_t *>(p) + w.lane;
        const int sh = 16 * w.lane;
        h[0] = (uint16_t)(v.x >> sh);
        h[2] = (uint16_t)(v.y >> sh);
        h[4] = (uint16_t)(v.z >> sh);
        h[6] = (uint16_t)(v.w >> sh);
    }
}

// one row's (DIAG plane, UP plane) of a row pair's direction word: half 0 = even row (.x/.y), 1 = odd row (.z/.w)
template <bool SOLO>
__device__ __forceinline__ void store_half(uint4 *p, int half, uint2 v, const FastWork &w) {
    if (!SOLO) {
        reinterpret_cast<uint2 *>(p)[half] = v;
    } else {
        uint16_t *h = reinterpret_cast<uint16_t *>(p) + 4 * half + w.lane;
        const int sh = 16 * w.lane;
        h[0] = (uint16_t)(v.x >> sh);
        h[2] = (uint16_t)(v.y >> sh);
    }
}

// PRMT selectors of a FULL strip of the packed inter-task kernels (columns c0 .. c0 + TW - 1, all inside the ref):
// sel[k] pairs, in its low 16 bits, lane A's "table byte f, then its sign" with lane B's (bytes 4..7 of the PRMT
// sources).  The TW ref codes of each lane come in as three 16-byte chunks, are shifted to byte 0 as words (the strip
// starts at an arbitrary column of its first chunk), turned into selector bytes four columns at a time, and one PRMT per
// column pairs the two lanes' bytes -- 0.5 k instructions per strip where byte loads with their 64-bit addresses took
// 1.8 k, and the integer pipe is what bounds the score kernels and the tagged align kernel.
template <int TW>
__device__ __forceinline__ void fast_strip_selectors(const ChunkGeom &g, const uint4 *__restrict__ code_refs, int slot_a, int c0,
                                                     uint32_t (&sel)[TW]) {
    static_assert(TW <= 32, "three 16-byte chunks hold a strip that starts anywhere in the first");
    const int a = c0 & 15, q0 = c0 >> 4;
    uint32_t wa[12], wb[12];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int q = min(q0 + t, g.ref_chunks - 1);
        const uint4 *pc = code_refs + (size_t)q * g.slots + slot_a;
        const uint4 va = pc[0], vb = pc[1];
        wa[4 * t] = va.x, wa[4 * t + 1] = va.y, wa[4 * t + 2] = va.z, wa[4 * t + 3] = va.w;
        wb[4 * t] = vb.x, wb[4 * t + 1] = vb.y, wb[4 * t + 2] = vb.z, wb[4 * t + 3] = vb.w;
    }
    if (a & 8) {
#pragma unroll
        for (int i = 0; i < 10; ++i) wa[i] = wa[i + 2], wb[i] = wb[i + 2];
    }
    if (a & 4) {
#pragma unroll
        for (int i = 0; i < 9; ++i) wa[i] = wa[i + 1], wb[i] = wb[i + 1];
    }
    const int bsh = (a & 3) * 8;
    constexpr int NW4 = (TW + 3) / 4;
#pragma unroll
    for (int i = 0; i < NW4; ++i) {
        const uint32_t ca = __funnelshift_r(wa[i], wa[i + 1], bsh), cb = __funnelshift_r(wb[i], wb[i + 1], bsh);
        // per byte: lane A  f | (f | 8) << 4,  lane B  (f | 4) | (f | 12) << 4   (f <= 5: nothing crosses a byte)
        wa[i] = ca | ((ca | 0x08080808u) << 4);
        wb[i] = (cb | 0x04040404u) | ((cb | 0x0C0C0C0Cu) << 4);
    }
#pragma unroll
    for (int k = 0; k < TW; ++k) sel[k] = prmt(wa[k >> 2], wb[k >> 2], (uint32_t)((k & 3) | ((4 + (k & 3)) << 4)));
}

__device__ __forceinline__ int fast_groups(int tw) { return (tw + 15) >> 4; }

// Intra-task kernels (va_intra.cu): a lane's strip is 16 columns; direction words are uint2 (DIAG plane, second
// plane; low 16 bits lane A) at [duo][strip][row], rows padded to a multiple of 4 so that four rows are one
// aligned 32-byte sector.
constexpr int INTRA_TW = 16;
__host__ __device__ __forceinline__ int intra_dir_rows(const ChunkGeom &g) { return (g.rows_alloc + 3) & ~3; }

// Direction words of the packed kernel: one uint4 per (strip, ROW PAIR, group of 16 columns, duo)
//   .x/.y = DIAG plane / UP plane of the even row, .z/.w = the same for the odd row;
//   in each 32-bit plane the low half belongs to lane A (slot 2u), the high half to lane B.
// Two rows per word keep the fill kernel's stores 16 bytes wide and fully coalesced and halve
// the number of dependent loads of the traceback walk.
__device__ __forceinline__ int fast_row_pairs(const ChunkGeom &g) { return (g.rows_alloc + 1) >> 1; }

// Packed NW align, duo lanes of different read lengths: sweep row of a lane's matrix row r is r + this
__device__ __forceinline__ int nw_row_offset(const PairMeta &me, const PairMeta &other) { return max((int)me.rows, (int)other.rows) - (int)me.rows; }
__device__ __forceinline__ int nw_row_offset(const ChunkGeom &g, const PairMeta &me, const PairMeta &other) { return g.intra ? 0 : nw_row_offset(me, other); }

// index (in uint4 units) of the word holding `row`
__device__ __forceinline__ size_t fast_dir_index(const ChunkGeom &g, int strip, int row, int group, int duo) {
    return (((size_t)strip * fast_row_pairs(g) + (row >> 1)) * fast_groups(g.fast_tw) + group) * g.duos + duo;
}

}  // namespace va
#endif
