/* versalign_cuda.h -- C ABI of libCUDAKernel.so, the B200 (sm_100a) kernel plug-in for
 * versalignLib's batched DP hot path.
 *
 * Two boundaries live in the same shared object:
 *
 *  (1) The reference's plug-in boundary, unchanged -- what its driver dlsym()s
 *      (src/Kernels/default/DefaultKernel_dllexport.cpp:18-42, identical in the SSE, AVX
 *      and OpenCL libraries):
 *          AlignmentKernel* spawn_alignment_kernel();
 *          void delete_alignment_kernel(AlignmentKernel*);
 *          void set_parameters(AlignmentParameters*);
 *          void set_logger(AlignmentLogger*);
 *      plus the C++-linkage globals _parameters / _logger the interface headers declare
 *      (include/AlignmentParameters.h:21, include/AlignmentLogger.h:21).  Declared in
 *      versalign_plugin_abi.h; implemented in csrc/cuda_kernel_plugin.cpp on top of (2).
 *
 *  (2) The flat C ABI below: plain pointers and sizes, no C++ types, no torch types.
 *      Each entry point names the reference interface it stands in for.  This is what a
 *      non-C++ host (ctypes, cgo, JNI...) binds, and what the plug-in class itself calls.
 *
 * Conventions shared with the reference:
 *   - a batch is n pairs; every read buffer holds exactly read_length bytes and every ref
 *     buffer exactly ref_length bytes, '\0' padded, not NUL terminated;
 *   - opt & 0xF: 0 = Smith-Waterman, 1 = "Needleman-Wunsch" (AlignmentKernel.h:26-32);
 *     any other value is a silent no-op that leaves the outputs untouched
 *     (DefaultKernel.cpp:35-40) -- these functions then return VA_OK;
 *   - scores are 16-bit; results are exact wherever no DP cell leaves the int16 range.
 *
 * There is no CPU fallback: without a usable CUDA device va_cuda_create fails.
 */
#ifndef VERSALIGN_CUDA_H
#define VERSALIGN_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VA_CUDA_ABI_VERSION 2

/* status codes */
#define VA_OK 0
#define VA_ERR_ARG -1     /* bad argument                                   */
#define VA_ERR_DEVICE -2  /* no CUDA device / CUDA runtime error            */
#define VA_ERR_MEMORY -3  /* host or device allocation failed               */
#define VA_ERR_RANGE -4   /* lengths or scores outside the supported domain */

#define VA_OPT_SW 0
#define VA_OPT_NW 1
/* Affine-gap (Gotoh) variants -- NOT in the reference: the same modes with a gap of length L
 * costing gap_open + L * gap_read (in the read) / gap_open + L * gap_ref (in the ref); gap_open == 0 reproduces
 * VA_OPT_SW / VA_OPT_NW bit for bit (policy VA_POLICY_DEFAULT_OCL, the only pointer rule of the variant).  The
 * reference dispatches on opt & 0xF only (DefaultKernel.cpp:31-40), so the gap-open score (<= 0) travels in the bits
 * above that nibble and every entry point below takes the variant unchanged:
 *     opt = VA_OPT_SW_AFFINE | VA_OPT_GAP_OPEN(-5)                                                              */
#define VA_OPT_SW_AFFINE 2
#define VA_OPT_NW_AFFINE 3
#define VA_OPT_GAP_OPEN(g) ((int)(((unsigned)(-(g)) & 0xFFFFu) << 8))

/* Traceback pointer rule.  The reference's kernels disagree:
 *   VA_POLICY_DEFAULT_OCL  DefaultKernel.cpp:238-248,338-346 and alignment_kernels.cl:106-112,334-339
 *   VA_POLICY_SIMD         SSEKernel.cpp:366-379,646-659 and AVXKernel.cpp:321-334,510-523          */
#define VA_POLICY_DEFAULT_OCL 0
#define VA_POLICY_SIMD 1

typedef struct va_cuda_ctx va_cuda_ctx;

/* The four scoring keys of AlignmentParameters (CustomParameters.h:9-47). */
typedef struct va_cuda_scoring {
    int32_t match;    /* score_match    */
    int32_t mismatch; /* score_mismatch */
    int32_t gap_read; /* score_gap_read */
    int32_t gap_ref;  /* score_gap_ref  */
} va_cuda_scoring;

/* Per-call phase timings of the last host-buffer call on this context (seconds, wall
 * clock on the calling thread's pipeline; phases overlap, so they do not add up to
 * total).  kernel_ms is CUDA-event time of the device work summed over chunks and taken
 * as the max over devices. */
typedef struct va_cuda_timings {
    double total_s;
    double gather_s;   /* host: scattered/flat input -> pinned staging     */
    double scatter_s;  /* host: pinned results -> caller's arrays          */
    double kernel_ms;  /* device: prep + fill (+ traceback) kernels        */
    int64_t cells;     /* DP cells actually computed (sum of rows x cols)  */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int32_t chunks;
    int32_t launches;  /* kernels launched                                 */
    int32_t devices;
    int32_t reserved;
} va_cuda_timings;

int va_cuda_abi_version(void);

/* Message of the last failure on the calling thread ("" if none). */
const char *va_cuda_last_error(void);

/* Number of visible CUDA devices (0 and VA_ERR_DEVICE when there is none). */
int va_cuda_device_count(int *count);

/* Create a context on the given device ordinals (n_devices == 0: every visible device).
 * Stands in for the kernel constructors' environment set-up
 * (OpenCLKernel.cpp:311-326 initialize_opencl_environment).  host_threads: worker threads
 * for staging (<= 0: choose from the core count); the reference's num_threads key. */
int va_cuda_create(va_cuda_ctx **ctx, const int *devices, int n_devices, int host_threads);
void va_cuda_destroy(va_cuda_ctx *ctx);
int va_cuda_set_host_threads(va_cuda_ctx *ctx, int host_threads);
int va_cuda_get_timings(const va_cuda_ctx *ctx, va_cuda_timings *out);

/* ---- host buffers, scattered: the reference's own calling convention ---------------- */

/* AlignmentKernel::score_alignments (AlignmentKernel.h:40-41; DefaultKernel.cpp:52-81,
 * SSEKernel.cpp:132-224, OpenCLKernel.cpp:28-162).  reads[i] / refs[i] are independent
 * heap blocks; scores[0..n) is overwritten.  Pairs are sharded over the context's
 * devices, no inter-device exchange. */
int va_cuda_score_ptrs(va_cuda_ctx *ctx, int opt, const va_cuda_scoring *sc, int n,
                       const char *const *reads, int read_length,
                       const char *const *refs, int ref_length, int16_t *scores);

/* AlignmentKernel::compute_alignments (AlignmentKernel.h:42-43; DefaultKernel.cpp:21-50,
 * 391-525; SSEKernel.cpp:41-130,729-1005; OpenCLKernel.cpp:164-309).  out_read[i] /
 * out_ref[i] must each point at read_length+ref_length writable bytes (the plug-in class
 * passes fresh new char[] blocks).  On return bytes [start[i], L-1) hold the gapped
 * strings, byte L-1 is NUL, bytes before start[i] are left untouched; start[i] is the value
 * the reference stores in readStart and refStart (readEnd = refEnd = L-1).  end_cell may
 * be NULL, else receives 2 shorts per pair: 0-based (read_pos, ref_pos) traceback began at. */
int va_cuda_align_ptrs(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n,
                       const char *const *reads, int read_length,
                       const char *const *refs, int ref_length,
                       char *const *out_read, char *const *out_ref, int16_t *start, int16_t *end_cell);

/* Same, but the result blocks are obtained from the caller's allocator while results stream
 * back -- on the staging threads, overlapped with device work -- instead of being
 * pre-allocated: out_read[i] = alloc(L, user), out_ref[i] = alloc(L, user).  This is what the
 * plug-in class uses with alloc = new char[] (DefaultKernel.cpp:441-442).  alloc must be
 * thread-safe; a NULL return aborts the call with VA_ERR_MEMORY (blocks handed out so far stay
 * in out_read/out_ref, untouched entries are NULL). */
typedef char *(*va_cuda_alloc_fn)(size_t bytes, void *user);
int va_cuda_align_alloc(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n,
                        const char *const *reads, int read_length,
                        const char *const *refs, int ref_length,
                        va_cuda_alloc_fn alloc, void *user,
                        char **out_read, char **out_ref, int16_t *start, int16_t *end_cell);

/* Same again, but the results are written straight into the caller's array of result records
 * laid out like the reference's struct Alignment (include/AlignmentKernel.h:12-24): two block
 * pointers followed by four shorts.  records points at record 0, record_stride is sizeof the
 * caller's struct (>= sizeof(va_cuda_alignment_record)).  Per pair: read / ref = alloc(L, user),
 * read_start = ref_start = first used index, read_end = ref_end = L-1 (DefaultKernel.cpp:441-453).
 * The plug-in class calls this with its Alignment[n] array, so no per-pair pass is left on the
 * calling thread after the pipeline has drained. */
typedef struct va_cuda_alignment_record {
    char *read;
    char *ref;
    int16_t read_start, read_end, ref_start, ref_end;
} va_cuda_alignment_record;
int va_cuda_align_records(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n,
                          const char *const *reads, int read_length,
                          const char *const *refs, int ref_length,
                          va_cuda_alloc_fn alloc, void *user,
                          void *records, size_t record_stride, int16_t *end_cell);

/* ---- host buffers, contiguous (fixed stride) ----------------------------------------- */

/* Same as va_cuda_score_ptrs with reads = n*read_length contiguous bytes (the layout
 * OpenCLKernel.cpp:61-66 gathers into). */
int va_cuda_score_flat(va_cuda_ctx *ctx, int opt, const va_cuda_scoring *sc, int n,
                       const char *reads, int read_length, const char *refs, int ref_length,
                       int16_t *scores);

/* Same as va_cuda_align_ptrs with contiguous outputs: aln_read / aln_ref are n blocks of
 * L = read_length+ref_length bytes (the layout OpenCLKernel.cpp:585-611 uses); bytes before
 * start[i] are zero. */
int va_cuda_align_flat(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n,
                       const char *reads, int read_length, const char *refs, int ref_length,
                       char *aln_read, char *aln_ref, int16_t *start, int16_t *end_cell);

/* ---- batch-friendly entry points (beside the reference's convention) ------------------ */

/* The reference's interface forces three host-side costs on a large batch: every sequence is a
 * separate heap block padded to the batch maximum (versalignUtil.cpp:17-33 pad()), every result
 * is two more heap blocks of read_length+ref_length bytes (DefaultKernel.cpp:441-442), and the
 * gapped strings are ~4x the information of the alignment itself.  These two entry points keep
 * the kernels and their semantics (same scores, same end cells, same paths) and change only the
 * containers:
 *   in   read i = reads[read_off[i] .. read_off[i+1]), ref i likewise: contiguous ASCII bases,
 *        per-pair lengths (<= 32000 each, read+ref <= 32767), no padding, no terminator;
 *        n+1 offsets per side
 *   out  scores[n];  coords[n][4] = read_begin, read_end, ref_begin, ref_end of the aligned
 *        region in SEQUENCE coordinates (0-based, half open);  a CIGAR per pair in BAM encoding
 *        (length << 4 | op, op 0 = M read and ref base, 1 = I read base against a gap, 2 = D ref
 *        base against a gap), pair i's runs at (*cigar)[cigar_off[i] .. cigar_off[i+1]).
 * Results leave the device as 2-bit moves (~(read+ref)/4 bytes per pair instead of 2*(read+ref)).
 * *cigar is ONE block obtained from alloc(bytes, user) after the batch is done (the caller frees
 * it); pass cigar = NULL to skip CIGARs, coords / scores = NULL to skip those. */
int va_cuda_score_packed(va_cuda_ctx *ctx, int opt, const va_cuda_scoring *sc, int n,
                         const char *reads, const int64_t *read_off,
                         const char *refs, const int64_t *ref_off, int16_t *scores);
int va_cuda_align_packed(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n,
                         const char *reads, const int64_t *read_off,
                         const char *refs, const int64_t *ref_off,
                         int16_t *scores, int32_t *coords, int64_t *cigar_off,
                         va_cuda_alloc_fn alloc, void *user, uint32_t **cigar);

/* ---- device-resident buffers (device 0 of the context) ------------------------------- */

/* Inputs and outputs already in HBM, same flat layouts as above; work is enqueued on
 * `stream` (a cudaStream_t, NULL = the legacy default stream) and NOT synchronised.
 * This is the kernel-only path bench.py times with CUDA events. */
int va_cuda_score_device(va_cuda_ctx *ctx, int opt, const va_cuda_scoring *sc, int n,
                         const void *d_reads, int read_length, const void *d_refs, int ref_length,
                         void *d_scores, void *stream);
int va_cuda_align_device(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n,
                         const void *d_reads, int read_length, const void *d_refs, int ref_length,
                         void *d_aln_read, void *d_aln_ref, void *d_start, void *d_end_cell,
                         void *stream);

/* Largest n one device-resident call accepts for these lengths with the workspace the
 * context may allocate (direction matrix for align calls). */
int va_cuda_max_resident_pairs(va_cuda_ctx *ctx, int align, int read_length, int ref_length, int64_t *max_n);

/* Per-kernel device times of the LAST device-resident call, for roofline accounting:
 * when profiling is on, CUDA events are recorded on the caller's stream around the prep, fill
 * and traceback kernels; va_cuda_get_kernel_ms waits for them and returns the elapsed
 * milliseconds summed over the call's sub-chunks (ms[0] prep, ms[1] fill, ms[2] traceback). */
int va_cuda_set_profiling(va_cuda_ctx *ctx, int on);
int va_cuda_get_kernel_ms(va_cuda_ctx *ctx, float ms[3]);

/* Timings of the last host-buffer call made through the plug-in class (the context the
 * CUDAKernel instances share).  VA_ERR_ARG before the first spawn. */
int va_cuda_plugin_timings(va_cuda_timings *out);

/* Integer-pipe microbenchmark used for the roofline denominator: runs `iters` dependent
 * rounds of `chains` independent VIADDMNMX chains per thread on every SM and reports
 * lane-operations per second (kind 0: .S32, 1: .S16x2 counted as 2 lanes, 2: .S16x2.RELU,
 * 3: VIMNMX3.S16x2). */
int va_cuda_int_peak(va_cuda_ctx *ctx, int kind, double *lane_ops_per_s, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VERSALIGN_CUDA_H */
