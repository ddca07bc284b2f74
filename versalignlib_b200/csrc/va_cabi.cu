// va_cabi.cu -- the flat C ABI of include/versalign_cuda.h: contexts, per-device engines,
// pinned staging, chunked pipelines and multi-GPU sharding.  No DP arithmetic here.
//
// Structural counterpart in the reference: the host half of its OpenCL kernel
// (OpenCLKernel.cpp:28-309): size a batch to the device (:517-568), gather the scattered
// sequences into contiguous host memory (:61-66), run the device kernel per batch (:91-96),
// copy the results out (:613-645).  Differences that matter on a B200: staging buffers are
// pinned and reused, chunks are double/triple buffered on CUDA streams so gather, H2D,
// kernels, D2H and scatter overlap, and the pair range is sharded over all devices of the
// context with one host thread per device and no inter-device exchange.
#include "versalign_cuda.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "va_internal.h"

namespace {

using namespace va;
using Clock = std::chrono::steady_clock;

thread_local std::string g_last_error;

int set_error(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return set_error(_e == cudaErrorMemoryAllocation ? VA_ERR_MEMORY : VA_ERR_DEVICE,       \
                             "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

inline double seconds_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }
inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------
// host worker pool: parallel_for that several device threads may call at the same time
// ---------------------------------------------------------------------------------------
class WorkerPool {
public:
    explicit WorkerPool(int threads) { resize(threads); }
    ~WorkerPool() { stop(); }

    void resize(int threads) {
        stop();
        n_ = std::max(1, threads);
        quit_ = false;
        for (int i = 0; i < n_ - 1; ++i) workers_.emplace_back([this] { loop(); });
    }
    int size() const { return n_; }

    // fn(begin, end) over [0, n) in blocks of `grain`; the caller works too.
    void parallel_for(int64_t n, int64_t grain, const std::function<void(int64_t, int64_t)> &fn) {
        if (n <= 0) return;
        if (n_ == 1 || n <= grain) {
            fn(0, n);
            return;
        }
        auto job = std::make_shared<Job>();
        job->n = n;
        job->grain = grain;
        job->fn = &fn;
        job->blocks = (n + grain - 1) / grain;
        job->remaining = job->blocks;
        {
            std::lock_guard<std::mutex> lk(mu_);
            jobs_.push_back(job);
        }
        cv_.notify_all();
        run(*job);
        std::unique_lock<std::mutex> lk(job->mu);
        job->done_cv.wait(lk, [&] { return job->remaining.load() == 0; });
        std::lock_guard<std::mutex> lk2(mu_);
        jobs_.erase(std::remove(jobs_.begin(), jobs_.end(), job), jobs_.end());
    }

private:
    struct Job {
        int64_t n = 0, grain = 1, blocks = 0;
        const std::function<void(int64_t, int64_t)> *fn = nullptr;
        std::atomic<int64_t> next{0};
        std::atomic<int64_t> remaining{0};
        std::mutex mu;
        std::condition_variable done_cv;
    };
    void run(Job &job) {
        for (;;) {
            const int64_t blk = job.next.fetch_add(1);
            if (blk >= job.blocks) return;
            const int64_t b = blk * job.grain, e = std::min(job.n, b + job.grain);
            (*job.fn)(b, e);
            if (job.remaining.fetch_sub(1) == 1) {
                std::lock_guard<std::mutex> lk(job.mu);
                job.done_cv.notify_all();
            }
        }
    }
    void loop() {
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] {
                    if (quit_) return true;
                    for (auto &j : jobs_)
                        if (j->next.load() < j->blocks) return true;
                    return false;
                });
                if (quit_) return;
                for (auto &j : jobs_)
                    if (j->next.load() < j->blocks) {
                        job = j;
                        break;
                    }
            }
            if (job) run(*job);
        }
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
        workers_.clear();
    }
    int n_ = 1;
    bool quit_ = false;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<std::shared_ptr<Job>> jobs_;
    std::vector<std::thread> workers_;
};

// ---------------------------------------------------------------------------------------
// grow-only buffers
// ---------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return VA_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = round_up(bytes + bytes / 8, 1 << 20);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, round_up(bytes, 256));
            if (e != cudaSuccess) {
                cudaGetLastError();
                return set_error(VA_ERR_MEMORY, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            }
            cap = round_up(bytes, 256);
        } else {
            cap = want;
        }
        return VA_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    // write_combined: staging the CPU only ever writes (inputs on their way to the device)
    int reserve(size_t bytes, bool write_combined = false) {
        if (bytes <= cap) return VA_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = round_up(bytes + bytes / 8, 1 << 16);
        cudaError_t e = cudaHostAlloc(&p, want, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_error(VA_ERR_MEMORY, "cudaHostAlloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return VA_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// Device workspace of one chunk in flight.
struct ChunkSlot {
    DevBuf raw_reads, raw_refs, code_reads, code_refs, row_idx, solo_list, meta, pair_of, prep_scratch, boundary, dirs, hrow, queue, scores, end_cell, aln_read, aln_ref, start, moves;
    PinBuf h_reads, h_refs, h_scores, h_end_cell, h_aln_read, h_aln_ref, h_start, h_moves;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_done = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    // side stream of the fill phase: the leftover kernels (solo slots, general) run beside the duo kernel
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // what is currently in flight in this slot
    int64_t first = 0;
    int count = 0;
    bool busy = false;

    void release() {
        DevBuf *d[] = {&raw_reads, &raw_refs, &code_reads, &code_refs, &row_idx, &solo_list, &meta, &pair_of, &prep_scratch, &boundary, &dirs, &hrow, &queue, &scores, &end_cell, &aln_read, &aln_ref, &start, &moves};
        for (auto *b : d) b->release();
        PinBuf *h[] = {&h_reads, &h_refs, &h_scores, &h_end_cell, &h_aln_read, &h_aln_ref, &h_start, &h_moves};
        for (auto *b : h) b->release();
        if (ev_done) cudaEventDestroy(ev_done);
        if (ev_k0) cudaEventDestroy(ev_k0);
        if (ev_k1) cudaEventDestroy(ev_k1);
        if (stream) cudaStreamDestroy(stream);
        if (side) cudaStreamDestroy(side);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        ev_done = ev_k0 = ev_k1 = ev_fork = ev_join = nullptr;
        stream = side = nullptr;
    }
};

constexpr int kRing = 3;

struct Engine {
    int device = 0;
    int sm_count = 0;
    size_t total_mem = 0;
    ChunkSlot ring[kRing];
    ChunkSlot resident;  // workspace of the device-resident entry points (no pinned memory, caller's stream)
    unsigned long long *d_cells = nullptr;
    unsigned int *d_sink = nullptr;
    // optional per-kernel events of device-resident calls (va_cuda_set_profiling)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;  // 4 per sub-chunk: before prep, after prep, after fill, after traceback
    size_t prof_used = 0;

    int init(int dev) {
        device = dev;
        CUDA_TRY(cudaSetDevice(dev));
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
        sm_count = prop.multiProcessorCount;
        total_mem = prop.totalGlobalMem;
        for (auto &s : ring) {
            CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreate(&s.ev_k0));
            CUDA_TRY(cudaEventCreate(&s.ev_k1));
        }
        CUDA_TRY(cudaMalloc(&d_cells, sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(d_cells, 0, sizeof(unsigned long long)));
        CUDA_TRY(cudaMalloc(&d_sink, 64));
        return VA_OK;
    }
    void release() {
        cudaSetDevice(device);
        for (auto &s : ring) s.release();
        resident.release();
        if (d_cells) cudaFree(d_cells);
        if (d_sink) cudaFree(d_sink);
        for (auto ev : prof_events) cudaEventDestroy(ev);
        prof_events.clear();
        d_cells = nullptr;
        d_sink = nullptr;
    }
};

// bytes of device workspace one pair needs (without raw inputs / outputs)
struct Shape {
    int read_length, ref_length, L;
    int read_chunks, ref_chunks, segs, rows_alloc;
    bool align;
    bool moves = false;  // align results leave the device as 2-bit move queues (packed entry points), not as strings
    size_t queue_words() const { return traceback_queue_words(read_length, ref_length); }
    // direction bytes per matrix row and pair: the general and the packed kernel keep separate
    // regions because one chunk can hold pairs of both kinds
    size_t gen_dir_row_bytes() const { return (size_t)segs * 2; }
    size_t dir_row_bytes() const { return gen_dir_row_bytes() + fast_dirs_bytes_per_row_per_slot(ref_length); }
    size_t per_pair_workspace() const {
        size_t b = (size_t)(read_chunks + ref_chunks) * 32 + (size_t)read_chunks * 8 + 2 * sizeof(PairMeta) + 40 + (size_t)rows_alloc * 6;
        if (align) {
            b += dir_row_bytes() * (rows_alloc + 1) + (size_t)ref_length * 2 + 8;
            const size_t qw = traceback_queue_words(read_length, ref_length);
            if (traceback_needs_global_queue(read_length, ref_length)) b += qw * 4;
        }
        return b;
    }
    size_t per_pair_io() const {
        size_t b = (size_t)read_length + ref_length + 2 + 4;
        if (align) b += (moves ? (queue_words() + 1) * 4 : 2 * (size_t)L) + 2;
        return b;
    }
};

Shape make_shape(int read_length, int ref_length, bool align) {
    Shape s;
    s.read_length = read_length;
    s.ref_length = ref_length;
    s.L = read_length + ref_length;
    s.read_chunks = (read_length + 15) / 16;
    s.ref_chunks = (ref_length + 15) / 16;
    s.segs = std::max(1, (ref_length + 7) / 8);
    s.rows_alloc = std::max(1, read_length);
    s.align = align;
    return s;
}

int mode_of(int opt, bool align) {
    const int alg = opt & 0xF;
    if (alg == VA_OPT_SW) return align ? MODE_SW_ALIGN : MODE_SW_SCORE;
    if (alg == VA_OPT_NW) return align ? MODE_NW_ALIGN : MODE_NW_SCORE;
    return -1;
}

int check_domain(const va_cuda_scoring *sc, int read_length, int ref_length) {
    if (!sc) return set_error(VA_ERR_ARG, "scoring is null");
    if (read_length < 0 || ref_length < 0 || read_length > 32000 || ref_length > 32000 || read_length + ref_length > 32767)
        return set_error(VA_ERR_RANGE, "read_length=%d ref_length=%d: offsets must fit a short (Alignment fields are short)",
                         read_length, ref_length);
    const int v[4] = {sc->match, sc->mismatch, sc->gap_read, sc->gap_ref};
    for (int x : v)
        if (x < -32768 || x > 32767) return set_error(VA_ERR_RANGE, "scoring value %d does not fit a short", x);
    return VA_OK;
}

}  // namespace

struct va_cuda_ctx {
    std::vector<Engine> engines;
    WorkerPool *pool = nullptr;
    int host_threads = 1;
    va_cuda_timings timings{};
    std::mutex call_mu;  // one host-buffer call at a time per context (the reference's callers are single threaded)
};

namespace {

int reserve_slot(ChunkSlot &s, const Shape &sh, int cap_pairs, bool pinned) {
    const size_t slots = round_up((size_t)cap_pairs, 64);
    int rc;
    if ((rc = s.raw_reads.reserve((size_t)cap_pairs * sh.read_length + 16))) return rc;
    if ((rc = s.raw_refs.reserve((size_t)cap_pairs * sh.ref_length + 16))) return rc;
    if ((rc = s.code_reads.reserve(slots * sh.read_chunks * 16 + 16))) return rc;
    if ((rc = s.code_refs.reserve(slots * sh.ref_chunks * 16 + 16))) return rc;
    if ((rc = s.row_idx.reserve(slots / 2 * sh.read_chunks * 16 + 16))) return rc;
    if ((rc = s.solo_list.reserve(slots * 4 + 64))) return rc;  // [0]: count, list from +16
    if ((rc = s.meta.reserve(slots * sizeof(PairMeta)))) return rc;
    if ((rc = s.pair_of.reserve(slots * 4))) return rc;
    if ((rc = s.prep_scratch.reserve(prep_scratch_bytes((int)slots, sh.read_length, sh.ref_length)))) return rc;
    // general region [rows][slots] + packed region [rows][duos] (kept apart: one chunk can hold both kinds)
    if ((rc = s.boundary.reserve(slots * sh.rows_alloc * 6 + 512))) return rc;
    if ((rc = s.scores.reserve(slots * 2))) return rc;
    if ((rc = s.end_cell.reserve(slots * 4))) return rc;
    if (sh.align) {
        if ((rc = s.dirs.reserve(slots * sh.dir_row_bytes() * (sh.rows_alloc + 1) + 512))) return rc;
        if ((rc = s.hrow.reserve(slots / 2 * (size_t)round_up((size_t)std::max(sh.ref_length, 1), 4) * 4 + 64))) return rc;
        // traceback move queue: shared memory unless the sequences are long
        const size_t qw = traceback_queue_words(sh.read_length, sh.ref_length);
        if (traceback_needs_global_queue(sh.read_length, sh.ref_length) && (rc = s.queue.reserve(slots * qw * 4 + 64))) return rc;
        if (pinned) {
            if (sh.moves) {
                if ((rc = s.moves.reserve((size_t)cap_pairs * (sh.queue_words() + 1) * 4 + 16))) return rc;
            } else {
                if ((rc = s.aln_read.reserve((size_t)cap_pairs * sh.L + 16))) return rc;
                if ((rc = s.aln_ref.reserve((size_t)cap_pairs * sh.L + 16))) return rc;
            }
            if ((rc = s.start.reserve(slots * 2))) return rc;
        }
    }
    if (pinned) {
        static const bool wc = [] { const char *v = getenv("VERSALIGN_CUDA_WC"); return v && atoi(v) != 0; }();
        if ((rc = s.h_reads.reserve((size_t)cap_pairs * sh.read_length + 16, wc))) return rc;
        if ((rc = s.h_refs.reserve((size_t)cap_pairs * sh.ref_length + 16, wc))) return rc;
        if ((rc = s.h_scores.reserve(slots * 2))) return rc;
        if ((rc = s.h_end_cell.reserve(slots * 4))) return rc;
        if (sh.align) {
            if (sh.moves) {
                if ((rc = s.h_moves.reserve((size_t)cap_pairs * (sh.queue_words() + 1) * 4 + 16))) return rc;
            } else {
                if ((rc = s.h_aln_read.reserve((size_t)cap_pairs * sh.L + 16))) return rc;
                if ((rc = s.h_aln_ref.reserve((size_t)cap_pairs * sh.L + 16))) return rc;
            }
            if ((rc = s.h_start.reserve(slots * 2))) return rc;
        }
    }
    return VA_OK;
}

void fill_geom(ChunkGeom &g, const Shape &sh, int n) {
    g.n = n;
    g.slots = (int)round_up((size_t)std::max(n, 1), 64);
    g.read_length = sh.read_length;
    g.ref_length = sh.ref_length;
    g.read_chunks = sh.read_chunks;
    g.ref_chunks = sh.ref_chunks;
    g.rows_alloc = sh.rows_alloc;
    g.segs = sh.segs;
    g.duos = g.slots / 2;
    g.fast_tw = 0;
    g.solo = 0;
}

// Enqueue prep + fill (+ traceback) for n pairs whose raw bytes are at raw_reads/raw_refs on
// the device.  Outputs go to the given device pointers.  Returns kernels launched (or < 0).
int enqueue_device_work(Engine &e, ChunkSlot &ws, const Shape &sh, int mode, int policy, const Scoring &sc, int n,
                        const uint8_t *raw_reads, const uint8_t *raw_refs, int16_t *scores, int16_t *end_cell,
                        uint8_t *aln_read, uint8_t *aln_ref, int16_t *start, bool zero_prefix, cudaStream_t stream,
                        bool profile = false, uint32_t *moves_out = nullptr) {
    ChunkGeom g;
    fill_geom(g, sh, n);
    // VERSALIGN_CUDA_GENERAL_ONLY=1 keeps every pair on the 32-bit kernel (parity tests use it to
    // cover that kernel on inputs the packed kernels would otherwise take)
    static const bool general_only = [] { const char *v = getenv("VERSALIGN_CUDA_GENERAL_ONLY"); return v && atoi(v) != 0; }();
    if (!general_only && fast_scoring_ok(mode, policy, sc, sh.read_length, sh.ref_length)) g.fast_tw = fast_pick_tw(mode, sh.ref_length);
    ChunkBuffers b{};
    b.raw_reads = raw_reads;
    b.raw_refs = raw_refs;
    b.code_reads = (uint4 *)ws.code_reads.p;
    b.code_refs = (uint4 *)ws.code_refs.p;
    b.row_idx = (uint4 *)ws.row_idx.p;
    b.hrow = (uint32_t *)ws.hrow.p;
    b.solo_count = (int32_t *)ws.solo_list.p;
    b.solo_list = (int32_t *)ws.solo_list.p + 16;
    b.meta = (PairMeta *)ws.meta.p;
    b.pair_of = (int32_t *)ws.pair_of.p;
    b.boundary = (int32_t *)ws.boundary.p;
    b.fboundary = (uint32_t *)((char *)ws.boundary.p + round_up((size_t)g.slots * sh.rows_alloc * 4, 256));
    b.dirs = (uint16_t *)ws.dirs.p;
    b.fdirs = (uint4 *)((char *)ws.dirs.p + round_up((size_t)g.slots * sh.gen_dir_row_bytes() * sh.rows_alloc, 256));
    b.scores = scores;
    b.end_cell = end_cell;
    b.aln_read = aln_read;
    b.aln_ref = aln_ref;
    b.start = start;
    b.moves_out = moves_out;
    b.cell_count = e.d_cells;
    int launches = 0;
    cudaEvent_t *pe = nullptr;
    if (profile) {
        if (e.prof_used + 4 > e.prof_events.size()) {
            for (int k = 0; k < 4; ++k) {
                cudaEvent_t ev;
                cudaEventCreate(&ev);
                e.prof_events.push_back(ev);
            }
        }
        pe = &e.prof_events[e.prof_used];
        e.prof_used += 4;
        cudaEventRecord(pe[0], stream);
    }
    static const bool no_intra = [] { const char *v = getenv("VERSALIGN_CUDA_NO_INTRA"); return v && atoi(v) != 0; }();
    const bool intra = g.fast_tw && !no_intra && intra_preferred(mode, n, sh.read_length, sh.ref_length, e.sm_count);
    // the inter-task kernels also take single slots (odd leftovers of the bucketing); the intra-task kernel
    // leaves those to the general kernel.  Decided before prep: every kernel of the chunk reads it.
    static const bool no_solo = [] { const char *v = getenv("VERSALIGN_CUDA_NO_SOLO"); return v && atoi(v) != 0; }();
    g.solo = (g.fast_tw && !intra && !no_solo) ? 1 : 0;
    launches += launch_prep(g, b, mode, policy, sc, ws.prep_scratch.p, ws.prep_scratch.cap, stream);
    if (pe) cudaEventRecord(pe[1], stream);
    // The packed duo kernel, the solo kernel and the general kernel own disjoint slots: the general kernel
    // (usually with nothing to do) runs beside the packed ones on a side stream instead of after them.
    if (!ws.side) {
        cudaStreamCreateWithFlags(&ws.side, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&ws.ev_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ws.ev_join, cudaEventDisableTiming);
    }
    cudaEventRecord(ws.ev_fork, stream);
    cudaStreamWaitEvent(ws.side, ws.ev_fork, 0);
    launches += launch_fill_general(g, b, mode, policy, sc, ws.side);
    cudaEventRecord(ws.ev_join, ws.side);
    if (intra)
        launches += launch_fill_intra(g, b, mode, make_fast_consts(mode, sc), stream);
    else
        launches += launch_fill_fast(g, b, mode, sc, stream);
    cudaStreamWaitEvent(stream, ws.ev_join, 0);
    if (pe) cudaEventRecord(pe[2], stream);
    if (sh.align) {
        if (zero_prefix && !moves_out) {  // bytes before start[i] are promised to be zero on this path
            cudaMemsetAsync(aln_read, 0, (size_t)n * sh.L, stream);
            cudaMemsetAsync(aln_ref, 0, (size_t)n * sh.L, stream);
        }
        launches += launch_traceback(g, b, mode, sc, (uint32_t *)ws.queue.p, stream);
    }
    if (pe) cudaEventRecord(pe[3], stream);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return set_error(VA_ERR_DEVICE, "kernel launch failed: %s", cudaGetErrorString(err));
    return launches;
}

// CIGAR runs of one chunk of a packed align call
struct CigarPart {
    int64_t first = 0;
    size_t words = 0;
    std::unique_ptr<uint32_t[]> data;
};

// What the host-buffer entry points have in common.
struct HostCall {
    int mode = 0, policy = 0;
    Scoring sc{};
    Shape sh{};
    int n = 0;
    // inputs: either scattered or flat
    const char *const *reads_p = nullptr;
    const char *const *refs_p = nullptr;
    const char *reads_f = nullptr;
    const char *refs_f = nullptr;
    // outputs
    int16_t *scores = nullptr;
    char *const *out_read_p = nullptr;
    char *const *out_ref_p = nullptr;
    va_cuda_alloc_fn alloc = nullptr;  // when set, out_*_w receive freshly allocated blocks
    void *alloc_user = nullptr;
    char **out_read_w = nullptr;
    char **out_ref_w = nullptr;
    std::atomic<int> *alloc_failed = nullptr;
    char *records = nullptr;  // when set (with alloc): va_cuda_alignment_record at records + i * record_stride
    size_t record_stride = 0;
    char *out_read_f = nullptr;
    char *out_ref_f = nullptr;
    int16_t *start = nullptr;
    int16_t *end_cell = nullptr;
    // packed (offset-addressed) inputs and CIGAR outputs of the batch-friendly entry points
    const int64_t *read_off = nullptr;
    const int64_t *ref_off = nullptr;
    int32_t *coords = nullptr;                        // [n][4]: read_begin, read_end, ref_begin, ref_end (0-based, half open)
    int64_t *cigar_off = nullptr;                     // [n+1]; filled with per-pair op counts first, prefix-summed at the end
    std::vector<CigarPart> *cigar_parts = nullptr;  // one per chunk, any order
    std::mutex *cigar_mu = nullptr;
};

struct ShardStats {
    double gather_s = 0, scatter_s = 0, kernel_ms = 0;
    int64_t h2d = 0, d2h = 0;
    int chunks = 0, launches = 0;
    unsigned long long cells = 0;
    int rc = VA_OK;
    std::string err;
};

void gather_chunk(va_cuda_ctx *ctx, const HostCall &c, ChunkSlot &s, int64_t first, int count) {
    char *hr = (char *)s.h_reads.p, *hf = (char *)s.h_refs.p;
    const int RL = c.sh.read_length, FL = c.sh.ref_length;
    if (c.read_off) {
        // offset-addressed sequences -> the kernels' fixed-stride, '\0'-padded staging layout
        ctx->pool->parallel_for(count, 2048, [&](int64_t b, int64_t e) {
            // no sequence is longer than RL / FL, so a block whose bytes add up to (e-b)*RL holds full-length
            // sequences only and already IS the staging layout
            const bool full_r = c.read_off[first + e] - c.read_off[first + b] == (e - b) * (int64_t)RL;
            const bool full_f = c.ref_off[first + e] - c.ref_off[first + b] == (e - b) * (int64_t)FL;
            if (full_r) memcpy(hr + b * RL, c.reads_f + c.read_off[first + b], (size_t)(e - b) * RL);
            if (full_f) memcpy(hf + b * FL, c.refs_f + c.ref_off[first + b], (size_t)(e - b) * FL);
            if (full_r && full_f) return;
            for (int64_t i = b; i < e; ++i) {
                const int64_t r0 = c.read_off[first + i], r1 = c.read_off[first + i + 1];
                const int64_t f0 = c.ref_off[first + i], f1 = c.ref_off[first + i + 1];
                memcpy(hr + i * RL, c.reads_f + r0, (size_t)(r1 - r0));
                memset(hr + i * RL + (r1 - r0), 0, (size_t)(RL - (r1 - r0)));
                memcpy(hf + i * FL, c.refs_f + f0, (size_t)(f1 - f0));
                memset(hf + i * FL + (f1 - f0), 0, (size_t)(FL - (f1 - f0)));
            }
        });
    } else if (c.reads_f) {
        ctx->pool->parallel_for(count, 4096, [&](int64_t b, int64_t e) {
            memcpy(hr + b * RL, c.reads_f + (first + b) * RL, (size_t)(e - b) * RL);
            memcpy(hf + b * FL, c.refs_f + (first + b) * FL, (size_t)(e - b) * FL);
        });
    } else {
        ctx->pool->parallel_for(count, 2048, [&](int64_t b, int64_t e) {
            // the blocks are scattered over the caller's heap: ask for the ones a few pairs ahead now
            constexpr int AHEAD = 8;
            for (int64_t i = b; i < e; ++i) {
                if (i + AHEAD < e) {
                    const char *pr = c.reads_p[first + i + AHEAD], *pf = c.refs_p[first + i + AHEAD];
                    for (int o = 0; o < RL; o += 64) __builtin_prefetch(pr + o, 0, 0);
                    for (int o = 0; o < FL; o += 64) __builtin_prefetch(pf + o, 0, 0);
                }
                memcpy(hr + i * RL, c.reads_p[first + i], RL);
                memcpy(hf + i * FL, c.refs_p[first + i], FL);
            }
        });
    }
}

// Packed entry points: a pair's result arrives as its moves in walk order (last alignment column first),
// normally already run-length encoded by the traceback kernel, else as the raw 2-bit queue
// (va_traceback.cu).  Reversed it is the CIGAR (BAM encoding: length << 4 | op; M = 0 read and ref base,
// I = 1 read base against a gap, D = 2 ref base against a gap) and the aligned coordinate ranges.
void scatter_moves(va_cuda_ctx *ctx, const HostCall &c, ChunkSlot &s, int64_t first, int count) {
    const size_t qw = c.sh.queue_words() + 1;
    const int16_t *ec = (const int16_t *)s.h_end_cell.p;
    const int16_t *sc = (const int16_t *)s.h_scores.p;
    const uint32_t *mv = (const uint32_t *)s.h_moves.p;
    if (c.scores) memcpy(c.scores + first, sc, (size_t)count * sizeof(int16_t));
    if (c.end_cell) memcpy(c.end_cell + 2 * first, ec, (size_t)count * 2 * sizeof(int16_t));
    // forward replay of pair i: fn(op, run length) per CIGAR run; returns the number of runs
    auto replay = [&](int64_t i, auto &&fn) -> int {
        const uint32_t *q = mv + (size_t)i * qw;
        const uint32_t head = q[0];
        ++q;
        if (!(head & 0x80000000u)) {  // runs, last first
            for (int r = (int)head - 1; r >= 0; --r) fn((int)(q[r] & 15u), (int)(q[r] >> 4));
            return (int)head;
        }
        const int n_moves = (int)(head & 0x7FFFFFFFu);
        int runs = 0, cur = -1, len = 0;
        for (int t = n_moves - 1; t >= 0; --t) {
            const int code = (q[t >> 4] >> (2 * (t & 15))) & 3;
            const int op = code == DIR_DIAG ? 0 : code == DIR_UP ? 1 : 2;
            if (op == cur) {
                ++len;
            } else {
                if (len) { fn(cur, len); ++runs; }
                cur = op;
                len = 1;
            }
        }
        if (len) { fn(cur, len); ++runs; }
        return runs;
    };
    // pass 1: runs per pair (the header says it unless the pair came back as raw moves)
    std::unique_ptr<int64_t[]> offs(new int64_t[(size_t)count + 1]);
    offs[0] = 0;
    ctx->pool->parallel_for(count, 8192, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) {
            const uint32_t head = mv[(size_t)i * qw];
            const int runs = (head & 0x80000000u) ? replay(i, [](int, int) {}) : (int)head;
            offs[(size_t)i + 1] = runs;
            if (c.cigar_off) c.cigar_off[first + i + 1] = runs;
        }
    });
    for (int64_t i = 0; i < count; ++i) offs[(size_t)i + 1] += offs[(size_t)i];
    // pass 2: coordinates and the runs themselves
    std::unique_ptr<uint32_t[]> part;
    uint32_t *pout = nullptr;
    if (c.cigar_parts) {
        part.reset(new uint32_t[(size_t)offs[(size_t)count] + 1]);  // not zero filled: every word is written below
        pout = part.get();
    }
    ctx->pool->parallel_for(count, 4096, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) {
            int used_read = 0, used_ref = 0;
            uint32_t *out = pout ? pout + offs[(size_t)i] : nullptr;
            replay(i, [&](int op, int len) {
                if (op != 2) used_read += len;
                if (op != 1) used_ref += len;
                if (out) *out++ = ((uint32_t)len << 4) | (uint32_t)op;
            });
            if (c.coords) {
                int32_t *co = c.coords + 4 * (first + i);
                co[1] = (int32_t)ec[2 * i] + 1;
                co[0] = co[1] - used_read;
                co[3] = (int32_t)ec[2 * i + 1] + 1;
                co[2] = co[3] - used_ref;
            }
        }
    });
    if (!c.cigar_parts) return;
    std::lock_guard<std::mutex> lk(*c.cigar_mu);
    c.cigar_parts->emplace_back();
    CigarPart &cp = c.cigar_parts->back();
    cp.first = first;
    cp.words = (size_t)offs[(size_t)count];
    cp.data = std::move(part);
}

void scatter_chunk(va_cuda_ctx *ctx, const HostCall &c, ChunkSlot &s, int64_t first, int count) {
    const int L = c.sh.L;
    if (!c.sh.align) {
        memcpy(c.scores + first, s.h_scores.p, (size_t)count * sizeof(int16_t));
        return;
    }
    const int16_t *st = (const int16_t *)s.h_start.p;
    const int16_t *ec = (const int16_t *)s.h_end_cell.p;
    if (c.sh.moves) {
        scatter_moves(ctx, c, s, first, count);
        return;
    }
    const char *ha = (const char *)s.h_aln_read.p, *hb = (const char *)s.h_aln_ref.p;
    if (c.start) memcpy(c.start + first, st, (size_t)count * sizeof(int16_t));
    if (c.end_cell) memcpy(c.end_cell + 2 * first, ec, (size_t)count * 2 * sizeof(int16_t));
    if (c.out_read_f) {
        ctx->pool->parallel_for(count, 2048, [&](int64_t b, int64_t e) {
            for (int64_t i = b; i < e; ++i) {
                int s0 = st[i];
                if (s0 < 0) s0 = 0;
                if (s0 > L) s0 = L;
                char *da = c.out_read_f + (first + i) * L, *db = c.out_ref_f + (first + i) * L;
                memset(da, 0, (size_t)s0);
                memset(db, 0, (size_t)s0);
                memcpy(da + s0, ha + i * L + s0, (size_t)(L - s0));
                memcpy(db + s0, hb + i * L + s0, (size_t)(L - s0));
            }
        });
    } else if (c.alloc && c.records) {
        ctx->pool->parallel_for(count, 1024, [&](int64_t b, int64_t e) {
            for (int64_t i = b; i < e; ++i) {
                int s0 = st[i];
                if (s0 < 0) s0 = 0;
                if (s0 > L) s0 = L;
                char *da = c.alloc((size_t)(L > 0 ? L : 1), c.alloc_user);
                char *db = c.alloc((size_t)(L > 0 ? L : 1), c.alloc_user);
                va_cuda_alignment_record *rec = reinterpret_cast<va_cuda_alignment_record *>(c.records + (size_t)(first + i) * c.record_stride);
                rec->read = da;
                rec->ref = db;
                rec->read_start = rec->ref_start = st[i];
                rec->read_end = rec->ref_end = (int16_t)(L - 1);
                if (!da || !db) {
                    c.alloc_failed->store(1);
                    continue;
                }
                memcpy(da + s0, ha + i * L + s0, (size_t)(L - s0));
                memcpy(db + s0, hb + i * L + s0, (size_t)(L - s0));
            }
        });
    } else if (c.alloc) {
        ctx->pool->parallel_for(count, 1024, [&](int64_t b, int64_t e) {
            for (int64_t i = b; i < e; ++i) {
                int s0 = st[i];
                if (s0 < 0) s0 = 0;
                if (s0 > L) s0 = L;
                char *da = c.alloc((size_t)(L > 0 ? L : 1), c.alloc_user);
                char *db = c.alloc((size_t)(L > 0 ? L : 1), c.alloc_user);
                c.out_read_w[first + i] = da;
                c.out_ref_w[first + i] = db;
                if (!da || !db) {
                    c.alloc_failed->store(1);
                    continue;
                }
                memcpy(da + s0, ha + i * L + s0, (size_t)(L - s0));
                memcpy(db + s0, hb + i * L + s0, (size_t)(L - s0));
            }
        });
    } else {
        ctx->pool->parallel_for(count, 1024, [&](int64_t b, int64_t e) {
            for (int64_t i = b; i < e; ++i) {
                int s0 = st[i];
                if (s0 < 0) s0 = 0;
                if (s0 > L) s0 = L;
                memcpy(c.out_read_p[first + i] + s0, ha + i * L + s0, (size_t)(L - s0));
                memcpy(c.out_ref_p[first + i] + s0, hb + i * L + s0, (size_t)(L - s0));
            }
        });
    }
}

// One device's share [lo, hi) of the batch, chunked through the ring.
void run_shard(va_cuda_ctx *ctx, Engine &e, const HostCall &c, int64_t lo, int64_t hi, int chunk_pairs, ShardStats &st) {
    auto fail = [&](int rc) {
        st.rc = rc;
        st.err = g_last_error;
    };
    if (cudaSetDevice(e.device) != cudaSuccess) return fail(set_error(VA_ERR_DEVICE, "cudaSetDevice(%d) failed", e.device));
    for (auto &s : e.ring) {
        int rc = reserve_slot(s, c.sh, chunk_pairs, true);
        if (rc) return fail(rc);
        s.busy = false;
    }
    cudaMemsetAsync(e.d_cells, 0, sizeof(unsigned long long), e.ring[0].stream);
    cudaStreamSynchronize(e.ring[0].stream);

    const int RL = c.sh.read_length, FL = c.sh.ref_length, L = c.sh.L;
    auto drain = [&](ChunkSlot &s) -> int {
        if (!s.busy) return VA_OK;
        cudaError_t err = cudaEventSynchronize(s.ev_done);
        if (err != cudaSuccess) return set_error(VA_ERR_DEVICE, "device work failed: %s", cudaGetErrorString(err));
        float ms = 0;
        if (cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1) == cudaSuccess) st.kernel_ms += ms;
        auto t0 = Clock::now();
        scatter_chunk(ctx, c, s, s.first, s.count);
        st.scatter_s += seconds_since(t0);
        s.busy = false;
        return VA_OK;
    };

    int k = 0;
    for (int64_t first = lo; first < hi; first += chunk_pairs, ++k) {
        const int count = (int)std::min<int64_t>(chunk_pairs, hi - first);
        ChunkSlot &s = e.ring[k % kRing];
        int rc = drain(s);  // the slot's previous chunk must be out before its pinned buffers are reused
        if (rc) return fail(rc);
        auto t0 = Clock::now();
        gather_chunk(ctx, c, s, first, count);
        st.gather_s += seconds_since(t0);

        cudaMemcpyAsync(s.raw_reads.p, s.h_reads.p, (size_t)count * RL, cudaMemcpyHostToDevice, s.stream);
        cudaMemcpyAsync(s.raw_refs.p, s.h_refs.p, (size_t)count * FL, cudaMemcpyHostToDevice, s.stream);
        st.h2d += (int64_t)count * (RL + FL);
        cudaEventRecord(s.ev_k0, s.stream);
        const bool zero_prefix = false;  // the host scatter zero-fills prefixes of the flat outputs itself
        int launches = enqueue_device_work(e, s, c.sh, c.mode, c.policy, c.sc, count, (const uint8_t *)s.raw_reads.p,
                                           (const uint8_t *)s.raw_refs.p, (int16_t *)s.scores.p, (int16_t *)s.end_cell.p,
                                           (uint8_t *)s.aln_read.p, (uint8_t *)s.aln_ref.p, (int16_t *)s.start.p,
                                           zero_prefix, s.stream, false, c.sh.moves ? (uint32_t *)s.moves.p : nullptr);
        if (launches < 0) return fail(launches);
        st.launches += launches;
        cudaEventRecord(s.ev_k1, s.stream);
        if (!c.sh.align) {
            cudaMemcpyAsync(s.h_scores.p, s.scores.p, (size_t)count * 2, cudaMemcpyDeviceToHost, s.stream);
            st.d2h += (int64_t)count * 2;
        } else if (c.sh.moves) {
            const size_t qb = (c.sh.queue_words() + 1) * 4;
            cudaMemcpyAsync(s.h_moves.p, s.moves.p, (size_t)count * qb, cudaMemcpyDeviceToHost, s.stream);
            cudaMemcpyAsync(s.h_start.p, s.start.p, (size_t)count * 2, cudaMemcpyDeviceToHost, s.stream);
            cudaMemcpyAsync(s.h_end_cell.p, s.end_cell.p, (size_t)count * 4, cudaMemcpyDeviceToHost, s.stream);
            cudaMemcpyAsync(s.h_scores.p, s.scores.p, (size_t)count * 2, cudaMemcpyDeviceToHost, s.stream);
            st.d2h += (int64_t)count * (int64_t)(qb + 8);
        } else {
            cudaMemcpyAsync(s.h_aln_read.p, s.aln_read.p, (size_t)count * L, cudaMemcpyDeviceToHost, s.stream);
            cudaMemcpyAsync(s.h_aln_ref.p, s.aln_ref.p, (size_t)count * L, cudaMemcpyDeviceToHost, s.stream);
            cudaMemcpyAsync(s.h_start.p, s.start.p, (size_t)count * 2, cudaMemcpyDeviceToHost, s.stream);
            cudaMemcpyAsync(s.h_end_cell.p, s.end_cell.p, (size_t)count * 4, cudaMemcpyDeviceToHost, s.stream);
            st.d2h += (int64_t)count * (2 * L + 6);
        }
        cudaEventRecord(s.ev_done, s.stream);
        s.first = first;
        s.count = count;
        s.busy = true;
        st.chunks++;
        // overlap: while this chunk runs on the device, hand the oldest finished chunk back
        ChunkSlot &prev = e.ring[(k + 1) % kRing];
        if (k >= kRing - 1) {
            rc = drain(prev);
            if (rc) return fail(rc);
        }
    }
    for (int j = 0; j < kRing; ++j) {
        int rc = drain(e.ring[(k + j) % kRing]);
        if (rc) return fail(rc);
    }
    unsigned long long cells = 0;
    if (cudaMemcpy(&cells, e.d_cells, sizeof(cells), cudaMemcpyDeviceToHost) == cudaSuccess) st.cells = cells;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) fail(set_error(VA_ERR_DEVICE, "CUDA error after shard: %s", cudaGetErrorString(err)));
}

int pick_chunk_pairs(const Engine &e, const Shape &sh, int64_t shard_pairs) {
    // device workspace budget per ring slot: a quarter of a third of the card, at most 6 GiB
    const size_t budget = std::min<size_t>(e.total_mem / 12, (size_t)6 << 30);
    const size_t per_pair = sh.per_pair_workspace() + sh.per_pair_io();
    int64_t cap = (int64_t)std::max<size_t>(64, budget / std::max<size_t>(per_pair, 1));
    // pinned staging per slot at most ~512 MiB
    cap = std::min<int64_t>(cap, std::max<int64_t>(64, ((int64_t)512 << 20) / (int64_t)std::max<size_t>(sh.per_pair_io(), 1)));
    // aim for >= 8 chunks per shard so the pipeline overlaps, but keep chunks >= min_chunk pairs (a chunk
    // costs a fixed ~0.3 ms of launches, copies and hand-offs)
    static const int64_t min_chunk = [] { const char *v = getenv("VERSALIGN_CUDA_MIN_CHUNK"); return v && atoll(v) > 0 ? atoll(v) : 65536LL; }();
    int64_t want = std::max<int64_t>((shard_pairs + 7) / 8, min_chunk);
    want = std::min<int64_t>(want, cap);
    want = std::min<int64_t>(want, std::max<int64_t>(shard_pairs, 1));
    want = (int64_t)round_up((size_t)want, 64);
    return (int)std::min<int64_t>(want, (int64_t)1 << 30);
}

int run_host_call(va_cuda_ctx *ctx, HostCall &c) {
    std::lock_guard<std::mutex> lk(ctx->call_mu);
    auto t0 = Clock::now();
    va_cuda_timings t{};
    t.devices = (int)ctx->engines.size();
    if (c.n > 0) {
        const int nd = (int)ctx->engines.size();
        std::vector<ShardStats> stats(nd);
        std::vector<std::thread> threads;
        // Contiguous slices per device.  Fixed-stride input: equal pair counts (== equal padded cells).
        // Offset-addressed input knows every length: cut where the running sum of rows x cols crosses
        // k/nd of the total, so mixed-length batches load the devices evenly.
        std::vector<int64_t> cut((size_t)nd + 1);
        for (int d = 0; d <= nd; ++d) cut[(size_t)d] = (int64_t)c.n * d / nd;
        if (nd > 1 && c.read_off) {
            std::vector<double> csum((size_t)c.n + 1, 0.0);
            for (int i = 0; i < c.n; ++i)
                csum[(size_t)i + 1] = csum[(size_t)i] + (double)(c.read_off[i + 1] - c.read_off[i]) * (double)(c.ref_off[i + 1] - c.ref_off[i]);
            for (int d = 1; d < nd; ++d)
                cut[(size_t)d] = std::lower_bound(csum.begin(), csum.end(), csum.back() * d / nd) - csum.begin();
            for (int d = 1; d <= nd; ++d) cut[(size_t)d] = std::max(cut[(size_t)d], cut[(size_t)d - 1]);
        }
        for (int d = 0; d < nd; ++d) {
            const int64_t lo = cut[(size_t)d], hi = cut[(size_t)d + 1];
            if (hi <= lo) continue;
            const int chunk = pick_chunk_pairs(ctx->engines[d], c.sh, hi - lo);
            if (nd == 1) {
                run_shard(ctx, ctx->engines[d], c, lo, hi, chunk, stats[d]);
            } else {
                threads.emplace_back([=, &stats, &c] { run_shard(ctx, ctx->engines[d], c, lo, hi, chunk, stats[d]); });
            }
        }
        for (auto &th : threads) th.join();
        for (auto &s : stats) {
            if (s.rc != VA_OK) {
                g_last_error = s.err;
                return s.rc;
            }
            t.gather_s = std::max(t.gather_s, s.gather_s);
            t.scatter_s = std::max(t.scatter_s, s.scatter_s);
            t.kernel_ms = std::max(t.kernel_ms, s.kernel_ms);
            t.cells += (int64_t)s.cells;
            t.h2d_bytes += s.h2d;
            t.d2h_bytes += s.d2h;
            t.chunks += s.chunks;
            t.launches += s.launches;
        }
    }
    t.total_s = seconds_since(t0);
    ctx->timings = t;
    return VA_OK;
}

int prepare_call(va_cuda_ctx *ctx, HostCall &c, int opt, bool align, int policy, const va_cuda_scoring *sc, int n,
                 int read_length, int ref_length, bool *noop) {
    *noop = false;
    if (!ctx) return set_error(VA_ERR_ARG, "context is null");
    if (n < 0) return set_error(VA_ERR_ARG, "n < 0");
    int rc = check_domain(sc, read_length, ref_length);
    if (rc) return rc;
    if (align && policy != VA_POLICY_DEFAULT_OCL && policy != VA_POLICY_SIMD) return set_error(VA_ERR_ARG, "unknown traceback policy %d", policy);
    c.mode = mode_of(opt, align);
    if (c.mode < 0) {  // unsupported algorithm: silent no-op like DefaultKernel.cpp:35-40
        *noop = true;
        return VA_OK;
    }
    c.policy = policy;
    c.sc = Scoring{sc->match, sc->mismatch, sc->gap_read, sc->gap_ref};
    c.sh = make_shape(read_length, ref_length, align);
    c.n = n;
    return VA_OK;
}

}  // namespace

// =========================================================================================
// exported C ABI
// =========================================================================================
extern "C" {

int va_cuda_abi_version(void) { return VA_CUDA_ABI_VERSION; }

const char *va_cuda_last_error(void) { return g_last_error.c_str(); }

int va_cuda_device_count(int *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (count) *count = (e == cudaSuccess) ? n : 0;
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error(VA_ERR_DEVICE, "no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return VA_OK;
}

int va_cuda_create(va_cuda_ctx **out, const int *devices, int n_devices, int host_threads) {
    if (!out) return set_error(VA_ERR_ARG, "ctx out pointer is null");
    *out = nullptr;
    int visible = 0;
    int rc = va_cuda_device_count(&visible);
    if (rc) return rc;
    std::vector<int> devs;
    if (n_devices <= 0 || !devices) {
        for (int d = 0; d < visible; ++d) devs.push_back(d);
    } else {
        for (int i = 0; i < n_devices; ++i) {
            if (devices[i] < 0 || devices[i] >= visible) return set_error(VA_ERR_ARG, "device %d not visible (%d devices)", devices[i], visible);
            devs.push_back(devices[i]);
        }
    }
    va_cuda_ctx *ctx = new va_cuda_ctx();
    ctx->engines.resize(devs.size());
    for (size_t i = 0; i < devs.size(); ++i) {
        rc = ctx->engines[i].init(devs[i]);
        if (rc) {
            va_cuda_destroy(ctx);
            return rc;
        }
    }
    ctx->pool = new WorkerPool(1);
    va_cuda_set_host_threads(ctx, host_threads);
    *out = ctx;
    return VA_OK;
}

void va_cuda_destroy(va_cuda_ctx *ctx) {
    if (!ctx) return;
    for (auto &e : ctx->engines) e.release();
    delete ctx->pool;
    delete ctx;
}

int va_cuda_set_host_threads(va_cuda_ctx *ctx, int host_threads) {
    if (!ctx) return set_error(VA_ERR_ARG, "context is null");
    int hw = (int)std::thread::hardware_concurrency();
    if (hw <= 0) hw = 1;
    int want = host_threads > 0 ? std::min(host_threads, hw) : std::min(hw, 32);
    if (want != ctx->host_threads || !ctx->pool) {
        std::lock_guard<std::mutex> lk(ctx->call_mu);
        ctx->pool->resize(want);
        ctx->host_threads = want;
    }
    return VA_OK;
}

int va_cuda_get_timings(const va_cuda_ctx *ctx, va_cuda_timings *out) {
    if (!ctx || !out) return set_error(VA_ERR_ARG, "null argument");
    *out = ctx->timings;
    return VA_OK;
}

int va_cuda_score_ptrs(va_cuda_ctx *ctx, int opt, const va_cuda_scoring *sc, int n, const char *const *reads,
                       int read_length, const char *const *refs, int ref_length, int16_t *scores) {
    HostCall c;
    bool noop;
    int rc = prepare_call(ctx, c, opt, false, 0, sc, n, read_length, ref_length, &noop);
    if (rc || noop) return rc;
    if (n > 0 && (!reads || !refs || !scores)) return set_error(VA_ERR_ARG, "null buffer");
    c.reads_p = reads;
    c.refs_p = refs;
    c.scores = scores;
    return run_host_call(ctx, c);
}

int va_cuda_score_flat(va_cuda_ctx *ctx, int opt, const va_cuda_scoring *sc, int n, const char *reads, int read_length,
                       const char *refs, int ref_length, int16_t *scores) {
    HostCall c;
    bool noop;
    int rc = prepare_call(ctx, c, opt, false, 0, sc, n, read_length, ref_length, &noop);
    if (rc || noop) return rc;
    if (n > 0 && (!reads || !refs || !scores)) return set_error(VA_ERR_ARG, "null buffer");
    c.reads_f = reads;
    c.refs_f = refs;
    c.scores = scores;
    return run_host_call(ctx, c);
}

int va_cuda_align_ptrs(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n, const char *const *reads,
                       int read_length, const char *const *refs, int ref_length, char *const *out_read,
                       char *const *out_ref, int16_t *start, int16_t *end_cell) {
    HostCall c;
    bool noop;
    int rc = prepare_call(ctx, c, opt, true, policy, sc, n, read_length, ref_length, &noop);
    if (rc || noop) return rc;
    if (n > 0 && (!reads || !refs || !out_read || !out_ref || !start)) return set_error(VA_ERR_ARG, "null buffer");
    c.reads_p = reads;
    c.refs_p = refs;
    c.out_read_p = out_read;
    c.out_ref_p = out_ref;
    c.start = start;
    c.end_cell = end_cell;
    return run_host_call(ctx, c);
}

int va_cuda_align_alloc(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n, const char *const *reads,
                        int read_length, const char *const *refs, int ref_length, va_cuda_alloc_fn alloc, void *user,
                        char **out_read, char **out_ref, int16_t *start, int16_t *end_cell) {
    HostCall c;
    bool noop;
    int rc = prepare_call(ctx, c, opt, true, policy, sc, n, read_length, ref_length, &noop);
    if (rc || noop) return rc;
    if (!alloc) return set_error(VA_ERR_ARG, "alloc is null");
    if (n > 0 && (!reads || !refs || !out_read || !out_ref || !start)) return set_error(VA_ERR_ARG, "null buffer");
    std::atomic<int> failed{0};
    for (int i = 0; i < n; ++i) out_read[i] = out_ref[i] = nullptr;
    c.reads_p = reads;
    c.refs_p = refs;
    c.alloc = alloc;
    c.alloc_user = user;
    c.out_read_w = out_read;
    c.out_ref_w = out_ref;
    c.alloc_failed = &failed;
    c.start = start;
    c.end_cell = end_cell;
    rc = run_host_call(ctx, c);
    if (rc == VA_OK && failed.load()) return set_error(VA_ERR_MEMORY, "the caller's allocator returned NULL");
    return rc;
}

int va_cuda_align_records(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n, const char *const *reads,
                          int read_length, const char *const *refs, int ref_length, va_cuda_alloc_fn alloc, void *user,
                          void *records, size_t record_stride, int16_t *end_cell) {
    HostCall c;
    bool noop;
    int rc = prepare_call(ctx, c, opt, true, policy, sc, n, read_length, ref_length, &noop);
    if (rc || noop) return rc;
    if (!alloc) return set_error(VA_ERR_ARG, "alloc is null");
    if (n > 0 && (!reads || !refs || !records)) return set_error(VA_ERR_ARG, "null buffer");
    if (record_stride < sizeof(va_cuda_alignment_record)) return set_error(VA_ERR_ARG, "record_stride is smaller than the record");
    std::atomic<int> failed{0};
    c.reads_p = reads;
    c.refs_p = refs;
    c.alloc = alloc;
    c.alloc_user = user;
    c.records = (char *)records;
    c.record_stride = record_stride;
    c.alloc_failed = &failed;
    c.end_cell = end_cell;
    rc = run_host_call(ctx, c);
    if (rc == VA_OK && failed.load()) return set_error(VA_ERR_MEMORY, "the caller's allocator returned NULL");
    return rc;
}

int va_cuda_align_flat(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n, const char *reads,
                       int read_length, const char *refs, int ref_length, char *aln_read, char *aln_ref, int16_t *start,
                       int16_t *end_cell) {
    HostCall c;
    bool noop;
    int rc = prepare_call(ctx, c, opt, true, policy, sc, n, read_length, ref_length, &noop);
    if (rc || noop) return rc;
    if (n > 0 && (!reads || !refs || !aln_read || !aln_ref || !start)) return set_error(VA_ERR_ARG, "null buffer");
    c.reads_f = reads;
    c.refs_f = refs;
    c.out_read_f = aln_read;
    c.out_ref_f = aln_ref;
    c.start = start;
    c.end_cell = end_cell;
    return run_host_call(ctx, c);
}

// ---- batch-friendly entry points: offset-addressed sequences in, scores / coordinates / CIGARs out ----
static int packed_lengths(int n, const int64_t *read_off, const int64_t *ref_off, int *read_length, int *ref_length) {
    int64_t rl = 0, fl = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t a = read_off[i + 1] - read_off[i], b = ref_off[i + 1] - ref_off[i];
        if (a < 0 || b < 0) return set_error(VA_ERR_ARG, "offsets of pair %d decrease", i);
        rl = std::max(rl, a);
        fl = std::max(fl, b);
    }
    if (rl > 32000 || fl > 32000) return set_error(VA_ERR_RANGE, "a sequence is longer than 32000 bases");
    *read_length = (int)rl;
    *ref_length = (int)fl;
    return VA_OK;
}

int va_cuda_score_packed(va_cuda_ctx *ctx, int opt, const va_cuda_scoring *sc, int n, const char *reads,
                         const int64_t *read_off, const char *refs, const int64_t *ref_off, int16_t *scores) {
    if (n < 0) return set_error(VA_ERR_ARG, "n < 0");
    if (n > 0 && (!reads || !refs || !read_off || !ref_off || !scores)) return set_error(VA_ERR_ARG, "null buffer");
    int rl = 0, fl = 0;
    int rc = packed_lengths(n, read_off, ref_off, &rl, &fl);
    if (rc) return rc;
    HostCall c;
    bool noop;
    rc = prepare_call(ctx, c, opt, false, 0, sc, n, rl, fl, &noop);
    if (rc || noop) return rc;
    c.reads_f = reads;
    c.refs_f = refs;
    c.read_off = read_off;
    c.ref_off = ref_off;
    c.scores = scores;
    return run_host_call(ctx, c);
}

int va_cuda_align_packed(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n, const char *reads,
                         const int64_t *read_off, const char *refs, const int64_t *ref_off, int16_t *scores,
                         int32_t *coords, int64_t *cigar_off, va_cuda_alloc_fn alloc, void *user, uint32_t **cigar) {
    if (n < 0) return set_error(VA_ERR_ARG, "n < 0");
    if (n > 0 && (!reads || !refs || !read_off || !ref_off)) return set_error(VA_ERR_ARG, "null buffer");
    if (cigar && (!alloc || !cigar_off)) return set_error(VA_ERR_ARG, "cigar output needs alloc and cigar_off");
    if (cigar) *cigar = nullptr;
    int rl = 0, fl = 0;
    int rc = packed_lengths(n, read_off, ref_off, &rl, &fl);
    if (rc) return rc;
    HostCall c;
    bool noop;
    rc = prepare_call(ctx, c, opt, true, policy, sc, n, rl, fl, &noop);
    if (rc || noop) return rc;
    c.sh.moves = true;
    c.reads_f = reads;
    c.refs_f = refs;
    c.read_off = read_off;
    c.ref_off = ref_off;
    c.scores = scores;
    c.coords = coords;
    c.cigar_off = cigar_off;
    std::vector<CigarPart> parts;
    std::mutex mu;
    if (cigar) {
        c.cigar_parts = &parts;
        c.cigar_mu = &mu;
    }
    if (cigar_off) cigar_off[0] = 0;
    rc = run_host_call(ctx, c);
    if (rc) return rc;
    if (cigar_off) {
        for (int i = 0; i < n; ++i) cigar_off[i + 1] += cigar_off[i];  // per-pair counts -> offsets
    }
    if (cigar) {
        const int64_t total = cigar_off[n];
        uint32_t *out = (uint32_t *)alloc((size_t)std::max<int64_t>(total, 1) * sizeof(uint32_t), user);
        if (!out) return set_error(VA_ERR_MEMORY, "the caller's allocator returned NULL");
        ctx->pool->parallel_for((int64_t)parts.size(), 1, [&](int64_t b, int64_t e) {
            for (int64_t k = b; k < e; ++k)
                if (parts[k].words) memcpy(out + cigar_off[parts[k].first], parts[k].data.get(), parts[k].words * sizeof(uint32_t));
        });
        *cigar = out;
    }
    return VA_OK;
}

int va_cuda_max_resident_pairs(va_cuda_ctx *ctx, int align, int read_length, int ref_length, int64_t *max_n) {
    if (!ctx || !max_n) return set_error(VA_ERR_ARG, "null argument");
    const Shape sh = make_shape(read_length, ref_length, align != 0);
    const Engine &e = ctx->engines[0];
    const size_t budget = std::min<size_t>(e.total_mem / 2, (size_t)80 << 30);
    *max_n = (int64_t)(budget / std::max<size_t>(sh.per_pair_workspace(), 1));
    return VA_OK;
}

static int resident_call(va_cuda_ctx *ctx, int opt, bool align, int policy, const va_cuda_scoring *sc, int n,
                         const void *d_reads, int read_length, const void *d_refs, int ref_length, void *d_scores,
                         void *d_aln_read, void *d_aln_ref, void *d_start, void *d_end_cell, void *stream) {
    HostCall c;
    bool noop;
    int rc = prepare_call(ctx, c, opt, align, policy, sc, n, read_length, ref_length, &noop);
    if (rc || noop) return rc;
    if (n == 0) return VA_OK;
    if (!d_reads || !d_refs) return set_error(VA_ERR_ARG, "null device buffer");
    Engine &e = ctx->engines[0];
    CUDA_TRY(cudaSetDevice(e.device));
    int64_t max_n = 0;
    va_cuda_max_resident_pairs(ctx, align, read_length, ref_length, &max_n);
    const int sub = (int)std::min<int64_t>(n, std::max<int64_t>(64, max_n / 64 * 64));
    rc = reserve_slot(e.resident, c.sh, sub, false);
    if (rc) return rc;
    ChunkSlot &ws = e.resident;
    // outputs the caller did not ask for still need somewhere to go
    if (!d_scores) {
        if ((rc = ws.scores.reserve(round_up((size_t)n, 64) * 2))) return rc;
    }
    if (align && !d_end_cell) {
        if ((rc = ws.end_cell.reserve(round_up((size_t)n, 64) * 4))) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int L = c.sh.L;
    int launches = 0;
    e.prof_used = 0;
    for (int64_t first = 0; first < n; first += sub) {
        const int count = (int)std::min<int64_t>(sub, n - first);
        int16_t *scores = d_scores ? (int16_t *)d_scores + first : (int16_t *)ws.scores.p + first;
        int16_t *endc = d_end_cell ? (int16_t *)d_end_cell + 2 * first : (int16_t *)ws.end_cell.p + (align ? 2 * first : 0);
        int k = enqueue_device_work(e, ws, c.sh, c.mode, c.policy, c.sc, count,
                                    (const uint8_t *)d_reads + first * read_length,
                                    (const uint8_t *)d_refs + first * ref_length, scores, endc,
                                    align ? (uint8_t *)d_aln_read + first * L : nullptr,
                                    align ? (uint8_t *)d_aln_ref + first * L : nullptr,
                                    align ? (int16_t *)d_start + first : nullptr, true, st, e.profiling);
        if (k < 0) return k;
        launches += k;
    }
    ctx->timings.launches = launches;
    return VA_OK;
}

int va_cuda_score_device(va_cuda_ctx *ctx, int opt, const va_cuda_scoring *sc, int n, const void *d_reads,
                         int read_length, const void *d_refs, int ref_length, void *d_scores, void *stream) {
    if (n > 0 && !d_scores) return set_error(VA_ERR_ARG, "d_scores is null");
    return resident_call(ctx, opt, false, 0, sc, n, d_reads, read_length, d_refs, ref_length, d_scores, nullptr, nullptr,
                         nullptr, nullptr, stream);
}

int va_cuda_align_device(va_cuda_ctx *ctx, int opt, int policy, const va_cuda_scoring *sc, int n, const void *d_reads,
                         int read_length, const void *d_refs, int ref_length, void *d_aln_read, void *d_aln_ref,
                         void *d_start, void *d_end_cell, void *stream) {
    if (n > 0 && (!d_aln_read || !d_aln_ref || !d_start)) return set_error(VA_ERR_ARG, "null device output buffer");
    return resident_call(ctx, opt, true, policy, sc, n, d_reads, read_length, d_refs, ref_length, nullptr, d_aln_read,
                         d_aln_ref, d_start, d_end_cell, stream);
}

int va_cuda_set_profiling(va_cuda_ctx *ctx, int on) {
    if (!ctx) return set_error(VA_ERR_ARG, "context is null");
    ctx->engines[0].profiling = on != 0;
    ctx->engines[0].prof_used = 0;
    return VA_OK;
}

int va_cuda_get_kernel_ms(va_cuda_ctx *ctx, float ms[3]) {
    if (!ctx || !ms) return set_error(VA_ERR_ARG, "null argument");
    Engine &e = ctx->engines[0];
    ms[0] = ms[1] = ms[2] = 0.f;
    for (size_t k = 0; k + 4 <= e.prof_used; k += 4) {
        CUDA_TRY(cudaEventSynchronize(e.prof_events[k + 3]));
        for (int j = 0; j < 3; ++j) {
            float t = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&t, e.prof_events[k + j], e.prof_events[k + j + 1]));
            ms[j] += t;
        }
    }
    return VA_OK;
}

int va_cuda_int_peak(va_cuda_ctx *ctx, int kind, double *lane_ops_per_s, void *stream) {
    if (!ctx || !lane_ops_per_s) return set_error(VA_ERR_ARG, "null argument");
    if (kind < 0 || kind > 3) return set_error(VA_ERR_ARG, "kind must be 0..3");
    Engine &e = ctx->engines[0];
    CUDA_TRY(cudaSetDevice(e.device));
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    double ops = 0, best = 0;
    launch_int_peak(kind, e.sm_count, 200, e.d_sink, st, &ops);  // warm-up
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a, st);
        launch_int_peak(kind, e.sm_count, 4000, e.d_sink, st, &ops);
        cudaEventRecord(b, st);
        CUDA_TRY(cudaEventSynchronize(b));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
        if (ms > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *lane_ops_per_s = best;
    return VA_OK;
}

}  // extern "C"
