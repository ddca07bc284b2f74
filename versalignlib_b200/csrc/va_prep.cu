// va_prep.cu -- staging on the device: raw bytes -> per-pair extents -> length-bucketed slot order
// -> base codes in slot order.
//
// What it replaces in the reference: the per-cell char_to_score[] look-ups
// (DefaultKernel.h:43-60 == opencl_definitions.cl:25-42) and the first-invalid-character scans of
// the NW fill (DefaultKernel.cpp:308-310,348-350; SSEKernel.cpp:514-518,673-677) are hoisted out
// of the DP loop; and where the reference's OpenCL host cuts the batch into equal pieces in input
// order (OpenCLKernel.cpp:517-568), pairs are first sorted by the DP extents they need, so that
// the threads of a warp -- and the two lanes of a packed pair-of-pairs -- sweep the same number of
// rows and columns on mixed-length batches.
//
//   meta_kernel    one thread per pair: word loads, table look-up per byte -> codes (pair order),
//                  last ACGT base / first invalid byte of each sequence -> PairMeta + sort key
//   radix sort     (cub) key = [dirty ref | cols | rows], value = pair index; stable, so a uniform
//                  batch keeps its order
//   encode_kernel  pair = order[slot]: codes and meta move to slot order, 16 bytes per thread
#include <algorithm>

#include <cub/device/device_radix_sort.cuh>

#include "va_fast.cuh"

namespace va {

namespace {

// byte -> code table (DefaultKernel.h:43-60 regrouped: A,C,G,T = 0..3, N = 4, everything else 5)
__device__ __forceinline__ void fill_lut(uint8_t *lut) {
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        const int u = c & 0xDF;  // fold case; bytes >= 0x80 keep bit 7 and stay OTHER
        lut[c] = (uint8_t)(u == 'A' ? CODE_A : u == 'C' ? CODE_C : u == 'G' ? CODE_G : u == 'T' ? CODE_T : u == 'N' ? CODE_N : CODE_OTHER);
    }
}

struct SeqScan {
    int last_acgt;      // index of the last ACGT base, -1 if none
    int first_other;    // first byte that is neither ACGT nor N (Default/OpenCL "invalid"), L if none
    int first_nonacgt;  // first byte that is not ACGT (SSE/AVX "invalid"), L if none
    int n_acgt;
};

// One thread translates and scans one sequence, 16 bases per step: aligned 32-bit loads (the
// sequence starts at an arbitrary byte, so neighbouring words are funnel-shifted into place), one
// shared-memory look-up per byte, and one coalesced uint4 store of 16 codes to the pair-interleaved
// scratch array codes[chunk][pair].  Bytes past L become OTHER.
__device__ __forceinline__ SeqScan scan_sequence(const uint8_t *__restrict__ raw, int L, int chunks, uint4 *__restrict__ codes,
                                                 size_t chunk_stride, const uint8_t *lut, const uint8_t *buf_end) {
    SeqScan s{-1, L, L, 0};
    const uintptr_t addr = reinterpret_cast<uintptr_t>(raw);
    const uint32_t *words = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
    const int shift = (int)(addr & 3) * 8;
    // Word wi of the sequence is funnel-shifted out of words[wi] and words[wi + 1]: the word path is taken only
    // while words[wi + 1] lies wholly inside the raw buffer (it ends at buf_end); the tail reads byte-wise.
    const long long whole = ((long long)(reinterpret_cast<uintptr_t>(buf_end) - reinterpret_cast<uintptr_t>(words)) >> 2) - 1;
    const int nwords_safe = (int)max(0LL, min((long long)((L + 3) / 4), whole));
    uint32_t carry = nwords_safe ? words[0] : 0;
    for (int c = 0; c < chunks; ++c) {
        uint32_t out[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int wi = c * 4 + q;  // sequence word: bytes 4*wi .. 4*wi+3
            uint32_t v;
            if (wi < nwords_safe) {
                const uint32_t next = words[wi + 1];
                v = __funnelshift_r(carry, next, shift);
                carry = next;
            } else {
                v = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int pos = wi * 4 + t;
                    if (pos < L) v |= (uint32_t)raw[pos] << (8 * t);
                }
            }
            uint32_t o = 0;
            if (wi * 4 + 4 <= L) {
                // whole word inside the sequence: four look-ups, then the statistics from bit masks
                // (codes 0..3 = ACGT have bit 2 clear; N = 4; OTHER = 5 has bits 2 and 0)
                o = (uint32_t)lut[v & 0xFF] | ((uint32_t)lut[(v >> 8) & 0xFF] << 8) | ((uint32_t)lut[(v >> 16) & 0xFF] << 16) |
                    ((uint32_t)lut[v >> 24] << 24);
                const uint32_t non = (o >> 2) & 0x01010101u;  // bit 0 of byte t: base t is not ACGT
                const uint32_t oth = non & o;                 // ... is neither ACGT nor N
                const uint32_t acgt = non ^ 0x01010101u;
                s.n_acgt += __popc(acgt);
                if (acgt) s.last_acgt = wi * 4 + ((31 - __clz(acgt)) >> 3);
                if (non && s.first_nonacgt == L) s.first_nonacgt = wi * 4 + ((__ffs(non) - 1) >> 3);
                if (oth && s.first_other == L) s.first_other = wi * 4 + ((__ffs(oth) - 1) >> 3);
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int pos = wi * 4 + t;
                    int code = CODE_OTHER;
                    if (pos < L) {
                        code = lut[(v >> (8 * t)) & 0xFF];
                        if (code < 4) {
                            s.last_acgt = pos;
                            s.n_acgt++;
                        } else {
                            s.first_nonacgt = min(s.first_nonacgt, pos);
                            if (code == CODE_OTHER) s.first_other = min(s.first_other, pos);
                        }
                    }
                    o |= (uint32_t)code << (8 * t);
                }
            }
            out[q] = o;
        }
        codes[(size_t)c * chunk_stride] = make_uint4(out[0], out[1], out[2], out[3]);
    }
    return s;
}

__global__ void __launch_bounds__(128) meta_kernel(ChunkGeom g, const uint8_t *__restrict__ raw_reads,
                                                   const uint8_t *__restrict__ raw_refs, const int64_t *__restrict__ read_off,
                                                   const int64_t *__restrict__ ref_off, PairMeta *__restrict__ meta_pair,
                                                   uint4 *__restrict__ codes_pair_reads, uint4 *__restrict__ codes_pair_refs,
                                                   uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, int mode,
                                                   int policy, int trim, int key_row_bits, int stage_bytes) {
    __shared__ uint8_t lut[256];
    extern __shared__ uint4 s_stage4[];  // staged raw bytes of the block's sequences (one side at a time), see below
    fill_lut(lut);
    __syncthreads();
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    // Short sequences: the block's 128 sequences of one side lie back to back in the raw buffer, so the block copies
    // that region into shared memory with coalesced 16-byte loads and every thread scans its sequence from there.
    // (Scanning straight from global memory, a warp's 4-byte loads are 150 bytes apart: 32 sectors per request, and
    // the kernel is bound by the L1's request rate -- 0.25 ms per million 150 bp pairs against 0.1 ms staged.)
    // `stage_bytes` = the dynamic shared memory the launcher gave the block; 0 or too small for this block: no staging.
    uint8_t *const s_stage = reinterpret_cast<uint8_t *>(s_stage4);
    auto staged_scan = [&](const uint8_t *raw, const int64_t *off, int stride, int seq_len, int chunks, uint4 *codes) -> SeqScan {
        const int first = blockIdx.x * blockDim.x, last = min(first + (int)blockDim.x, g.n);  // the block's pairs [first, last)
        const uint8_t *lo = seq_ptr(raw, off, first, stride), *hi = seq_ptr(raw, off, last, stride);  // (pair g.n: the end of the buffer)
        const uint8_t *buf_end = seq_end(raw, off, g.n, stride);
        const uintptr_t base = reinterpret_cast<uintptr_t>(lo) & ~(uintptr_t)15;
        const size_t span = (size_t)(reinterpret_cast<uintptr_t>(hi) - base);  // bytes from the aligned base to the region's end
        const bool staged = stage_bytes > 0 && span + 16 <= (size_t)stage_bytes;  // block-uniform
        if (staged) {
            const int n16 = (int)((span + 15) >> 4);
            for (int q = threadIdx.x; q < n16; q += blockDim.x) {
                const uint8_t *src = reinterpret_cast<const uint8_t *>(base) + (size_t)q * 16;
                if (src + 16 <= buf_end) {
                    s_stage4[q] = *reinterpret_cast<const uint4 *>(src);
                } else {  // the buffer's last, partial 16 bytes
                    for (int t = 0; t < 16; ++t) s_stage[(size_t)q * 16 + t] = src + t < buf_end ? src[t] : (uint8_t)0;
                }
            }
        }
        __syncthreads();
        SeqScan r{-1, seq_len, seq_len, 0};
        if (pair < g.n) {
            const uint8_t *mine = seq_ptr(raw, off, pair, stride);
            if (staged) {
                const uint8_t *sm = s_stage + (reinterpret_cast<uintptr_t>(mine) - base);
                r = scan_sequence(sm, seq_len, chunks, codes + pair, (size_t)g.slots, lut, s_stage + (((span + 15) >> 4) << 4));
            } else {
                r = scan_sequence(mine, seq_len, chunks, codes + pair, (size_t)g.slots, lut, buf_end);
            }
        }
        __syncthreads();  // the next side reuses the staging buffer
        return r;
    };
    {
        // offset-addressed sequences carry their own lengths; bytes past them count as the '\0' pad of the
        // fixed-stride layout (code OTHER)
        const int pc = min(pair, g.n - 1);
        const int read_len = read_off ? (int)(read_off[pc + 1] - read_off[pc]) : g.read_length;
        const int ref_len = ref_off ? (int)(ref_off[pc + 1] - ref_off[pc]) : g.ref_length;
        const SeqScan rd = staged_scan(raw_reads, read_off, g.read_length, read_len, g.read_chunks, codes_pair_reads);
        const SeqScan rf = staged_scan(raw_refs, ref_off, g.ref_length, ref_len, g.ref_chunks, codes_pair_refs);
        if (pair >= g.n) return;
        PairMeta m;
        m.true_rows = (int16_t)(rd.last_acgt + 1);
        m.true_cols = (int16_t)(rf.last_acgt + 1);
        m.flags = (int16_t)((rd.n_acgt != rd.last_acgt + 1 ? 1 : 0) | (rf.n_acgt != rf.last_acgt + 1 ? 2 : 0));
        m.max_read_pos = (int16_t)((policy == 1 ? rd.first_nonacgt : rd.first_other) - 1);
        m.max_ref_pos = (int16_t)((policy == 1 ? rf.first_nonacgt : rf.first_other) - 1);
        m.pad = 0;
        if (mode == MODE_NW_ALIGN) {
            // rows below the first invalid read character are never consulted; the end-cell rule scans
            // the whole padded width of the last valid row
            m.rows = (int16_t)(m.max_read_pos + 1);
            // With both gaps <= 0 the pad columns need not be filled: the fill/traceback kernels derive
            // whether the padded arg-max would land in them.  Trailing N of the ref stay inside the
            // sweep when the policy counts them as valid (max_ref_pos points past the last ACGT base).
            m.cols = trim ? (int16_t)max((int)m.true_cols, (int)m.max_ref_pos + 1) : (int16_t)g.ref_length;
        } else if (trim) {
            // trailing rows/columns that can only score 0 never change the result while both gap
            // scores are <= 0 (SURVEY.md A.1/A.2 "padding is neutral")
            m.rows = m.true_rows;
            m.cols = m.true_cols;
            if (mode == MODE_SW_ALIGN) {
                // with a zero score traceback starts at cell (0,0) (DefaultKernel.cpp:207-208): that
                // cell's pointer must exist even when a sequence holds no ACGT base at all
                m.rows = (int16_t)max((int)m.rows, min(1, g.read_length));
                m.cols = (int16_t)max((int)m.cols, min(1, g.ref_length));
            }
        } else {
            m.rows = (int16_t)g.read_length;
            m.cols = (int16_t)g.ref_length;
        }
        meta_pair[pair] = m;
        // pairs the packed kernels cannot take (padding / non-ACGT inside the swept ref columns) sort last
        const uint32_t dirty = ((m.flags & 2) || m.cols != m.true_cols) ? 1u : 0u;
        keys[pair] = (((dirty << 15) | (uint32_t)m.cols) << key_row_bits) | (uint32_t)m.rows;
        vals[pair] = (uint32_t)pair;
    }
}

// Pair order -> slot order: one thread moves one 16-base chunk (uint4) of one sequence into the
// slot-interleaved arrays; thread 0 of a slot also moves the meta record.
__global__ void __launch_bounds__(256) encode_kernel(ChunkGeom g, ChunkBuffers b, const PairMeta *__restrict__ meta_pair,
                                                     const uint4 *__restrict__ codes_pair_reads,
                                                     const uint4 *__restrict__ codes_pair_refs,
                                                     const uint32_t *__restrict__ order,
                                                     const uint32_t *__restrict__ sorted_keys, int mode) {
    // grid: x over the slots, y over the 16-base chunks (reads first, then refs) -- consecutive threads =
    // consecutive slots of the same chunk: coalesced 16-byte stores, no index division
    const int c = blockIdx.y;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < g.slots; slot += gridDim.x * blockDim.x) {
        if (slot >= g.n) {  // padding slots: only their meta is ever looked at (as a duo partner)
            if (c == 0) {
                PairMeta z{};
                b.meta[slot] = z;
                b.pair_of[slot] = -1;
            }
            continue;
        }
        const int pair = (int)order[slot];
        if (c == 0) {
            const PairMeta me = meta_pair[pair];
            b.meta[slot] = me;
            b.pair_of[slot] = pair;
            if (g.solo) {  // slots the packed kernels compute on their own (va_fast.cuh): compiled into a list
                PairMeta other{};
                if ((slot ^ 1) < g.n) other = meta_pair[order[slot ^ 1]];
                if (slot_owner(g, mode, slot, (slot & 1) ? other : me, (slot & 1) ? me : other) == OWN_SOLO)
                    b.solo_list[atomicAdd(b.solo_count, 1)] = slot;
            }
        }
        // scratch is [chunk][pair]: for a batch that keeps its order both sides are coalesced
        if (c < g.read_chunks) {
            const uint4 ca = codes_pair_reads[(size_t)c * g.slots + pair];
            b.code_reads[(size_t)c * g.slots + slot] = ca;
            if (b.row_idx && (slot & 1) == 0) {
                // the duo's row index: 7*code_a + code_b per byte (no carries: codes are <= 6)
                uint4 cb = make_uint4(0x05050505u, 0x05050505u, 0x05050505u, 0x05050505u);
                uint4 cx = ca;
                const bool has_b = slot + 1 < g.n;
                const int pair_b = has_b ? (int)order[slot + 1] : 0;
                if (has_b) cb = codes_pair_reads[(size_t)c * g.slots + pair_b];
                // (equal sort keys = equal extents: nothing to shift, and no need to look at the meta records)
                if (mode == MODE_NW_ALIGN && !g.intra && has_b && sorted_keys[slot] != sorted_keys[slot + 1]) {
                    // packed NW align end-aligns a duo's lanes: the lane with fewer rows starts late, behind
                    // CODE_PRE rows (va_internal.h); only duos the packed kernel takes are ever read back
                    // (slots that end up solo sweep from their own row 0: no shift for them)
                    const PairMeta ma = meta_pair[pair], mb = meta_pair[pair_b];
                    const bool duo = duo_is_fast(g, mode, slot, ma, mb);
                    const int off_a = duo ? nw_row_offset(ma, mb) : 0, off_b = duo ? nw_row_offset(mb, ma) : 0;
                    auto shifted = [&](int p, int off) {
                        uint32_t w[4];
                        const uint8_t *src = reinterpret_cast<const uint8_t *>(codes_pair_reads);
                        for (int r = 0; r < 16; ++r) {
                            const int row = c * 16 + r - off;
                            const uint32_t code = row < 0 ? (uint32_t)CODE_PRE : (uint32_t)src[((size_t)(row >> 4) * g.slots + p) * 16 + (row & 15)];
                            if ((r & 3) == 0) w[r >> 2] = 0;
                            w[r >> 2] |= code << (8 * (r & 3));
                        }
                        return make_uint4(w[0], w[1], w[2], w[3]);
                    };
                    if (off_a > 0) cx = shifted(pair, off_a);
                    if (off_b > 0) cb = shifted(pair_b, off_b);
                }
                b.row_idx[(size_t)c * g.duos + (slot >> 1)] = make_uint4(cx.x * 7u + cb.x, cx.y * 7u + cb.y, cx.z * 7u + cb.z, cx.w * 7u + cb.w);
            }
        } else if (c - g.read_chunks < g.ref_chunks) {
            b.code_refs[(size_t)(c - g.read_chunks) * g.slots + slot] = codes_pair_refs[(size_t)(c - g.read_chunks) * g.slots + pair];
        }
    }
}

int bits_for(int v) {
    int b = 1;
    while ((1 << b) <= v) ++b;
    return b;
}

}  // namespace

size_t prep_temp_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, n > 0 ? n : 1, 0, 32);
    return bytes + 256;
}

// scratch layout (all sized for `slots`): meta_pair | codes in pair order (reads, refs) | keys_in |
// keys_out | vals_in | vals_out | cub temp
size_t prep_scratch_bytes(int slots, int read_length, int ref_length) {
    const size_t rpad = (size_t)((read_length + 15) / 16) * 16, fpad = (size_t)((ref_length + 15) / 16) * 16;
    return (size_t)slots * (sizeof(PairMeta) + rpad + fpad + 4 * sizeof(uint32_t)) + prep_temp_bytes(slots) + 4096;
}

int launch_prep(const ChunkGeom &g, const ChunkBuffers &b, int mode, int policy, const Scoring &sc, void *scratch,
                size_t scratch_bytes, cudaStream_t stream) {
    if (g.n <= 0) return 0;
    auto align256 = [](char *q) { return (char *)(((uintptr_t)q + 255) & ~(uintptr_t)255); };
    char *p = (char *)scratch;
    PairMeta *meta_pair = (PairMeta *)p;
    p = align256(p + (size_t)g.slots * sizeof(PairMeta));
    uint4 *codes_reads = (uint4 *)p;
    p = align256(p + (size_t)g.slots * g.read_chunks * 16);
    uint4 *codes_refs = (uint4 *)p;
    p = align256(p + (size_t)g.slots * g.ref_chunks * 16);
    uint32_t *keys_in = (uint32_t *)p;
    p += (size_t)g.slots * 4;
    uint32_t *keys_out = (uint32_t *)p;
    p += (size_t)g.slots * 4;
    uint32_t *vals_in = (uint32_t *)p;
    p += (size_t)g.slots * 4;
    uint32_t *vals_out = (uint32_t *)p;
    p = align256(p + (size_t)g.slots * 4);
    size_t temp_bytes = scratch_bytes - (size_t)(p - (char *)scratch);

    const int trim = (sc.gap_read <= 0 && sc.gap_ref <= 0) ? 1 : 0;
    const int row_bits = bits_for(g.read_length);
    const int threads = 256;
    const int grid_cap = 148 * 8;
    const int meta_blocks = (g.n + 127) / 128;
    const dim3 enc_blocks((unsigned)std::min((g.slots + threads - 1) / threads, grid_cap * 4), (unsigned)std::max(1, g.read_chunks + g.ref_chunks));  // y >= 1: chunk 0 also moves the meta records
    // staging buffer of the meta kernel: the block's 128 sequences of one side + alignment slack, while that fits 40 KB
    const size_t stage_want = (size_t)128 * std::max(g.read_length, g.ref_length) + 48;
    const int stage_bytes = stage_want <= 40 * 1024 ? (int)((stage_want + 15) & ~(size_t)15) : 0;
    meta_kernel<<<meta_blocks, 128, stage_bytes, stream>>>(g, b.raw_reads, b.raw_refs, b.read_off, b.ref_off, meta_pair, codes_reads, codes_refs,
                                                             keys_in, vals_in, mode, policy, trim, row_bits, stage_bytes);
    // only the bits that can differ are sorted: rows, cols and the "dirty" flag above them
    cub::DeviceRadixSort::SortPairs(p, temp_bytes, keys_in, keys_out, vals_in, vals_out, g.n, 0, row_bits + 16, stream);
    if (g.solo) cudaMemsetAsync(b.solo_count, 0, sizeof(int32_t), stream);
    encode_kernel<<<enc_blocks, threads, 0, stream>>>(g, b, meta_pair, codes_reads, codes_refs,
                                                      vals_out, keys_out, mode);
    return 2;  // our kernels; the sort is the library's
}

}  // namespace va
