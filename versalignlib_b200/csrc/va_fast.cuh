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
//   - both lanes sweep the same columns; rows may differ for SW and the score modes (trailing rows
//     that can only score 0 are neutral) but not for NW align, whose end-cell rule reads one exact row.
__device__ __forceinline__ bool duo_is_fast(const ChunkGeom &g, int mode, int slot_a, const PairMeta &a, const PairMeta &b) {
    if (g.fast_tw == 0 || slot_a + 1 >= g.n) return false;
    if ((a.flags & 2) || (b.flags & 2)) return false;
    if (a.cols != b.cols || a.cols <= 0) return false;
    if (a.cols != a.true_cols || b.cols != b.true_cols) return false;  // no padded / N tail inside the sweep
    // NW align's end-cell rule reads the last valid row, which is taken from the registers after the
    // sweep: both lanes must end on the same row (the bucketing sort makes that the common case)
    if (mode == MODE_NW_ALIGN) return a.rows == b.rows && a.rows > 0;
    return max(a.rows, b.rows) > 0;
}

__device__ __forceinline__ int fast_groups(int tw) { return (tw + 15) >> 4; }

// Direction words of the packed kernel: one uint4 per (strip, ROW PAIR, group of 16 columns, duo)
//   .x/.y = DIAG plane / UP plane of the even row, .z/.w = the same for the odd row;
//   in each 32-bit plane the low half belongs to lane A (slot 2u), the high half to lane B.
// Two rows per word keep the fill kernel's stores 16 bytes wide and fully coalesced and halve
// the number of dependent loads of the traceback walk.
__device__ __forceinline__ int fast_row_pairs(const ChunkGeom &g) { return (g.rows_alloc + 1) >> 1; }

// index (in uint4 units) of the word holding `row`
__device__ __forceinline__ size_t fast_dir_index(const ChunkGeom &g, int strip, int row, int group, int duo) {
    return (((size_t)strip * fast_row_pairs(g) + (row >> 1)) * fast_groups(g.fast_tw) + group) * g.duos + duo;
}

}  // namespace va
#endif
