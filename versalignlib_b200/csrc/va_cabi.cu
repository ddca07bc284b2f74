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
#include "va_line_streamer.h"

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

// VERSALIGN_CUDA_TRACE=1: host-side timeline on stderr.  Call-level marks are ms since the entry point began.
bool trace_on() {
    static const bool on = [] { const char *v = getenv("VERSALIGN_CUDA_TRACE"); return v && atoi(v) != 0; }();
    return on;
}
thread_local Clock::time_point g_call_t0;
inline void call_begin() { if (trace_on()) g_call_t0 = Clock::now(); }
inline void call_mark(const char *what) {
    if (trace_on()) fprintf(stderr, "[va trace] call  %8.3f ms  %s\n", seconds_since(g_call_t0) * 1e3, what);
}
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
    DevBuf raw_reads, raw_refs, read_off, ref_off, code_reads, code_refs, row_idx, solo_list, meta, pair_of, prep_scratch, boundary, boundary_e, dirs, dirs4, zdirs, hrow, queue,
        scores, end_cell, start, moves, compact, compact_off, cursor, coords, run_count, run_offs, cigar, scan_tmp;
    PinBuf h_reads, h_refs, h_read_off, h_ref_off, h_scores, h_end_cell, h_start, h_compact, h_compact_off, h_cursor, h_coords, h_run_offs,
        h_cigar, h_moves;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_done = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    // side stream of the fill phase: the leftover kernels (solo slots, general) run beside the duo kernel
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // what is currently in flight in this slot
    int64_t first = 0;
    int count = 0;
    bool busy = false;
    size_t sent = 0;  // units of the variable-size result (string bytes / CIGAR runs) whose D2H copy is already enqueued

    void release() {
        DevBuf *d[] = {&raw_reads, &raw_refs, &read_off, &ref_off, &code_reads, &code_refs, &row_idx, &solo_list, &meta, &pair_of, &prep_scratch,
                       &boundary, &boundary_e, &dirs, &dirs4, &zdirs, &hrow, &queue, &scores, &end_cell, &start, &moves, &compact, &compact_off, &cursor, &coords,
                       &run_count, &run_offs, &cigar, &scan_tmp};
        for (auto *b : d) b->release();
        PinBuf *h[] = {&h_reads, &h_refs, &h_read_off, &h_ref_off, &h_scores, &h_end_cell, &h_start, &h_compact, &h_compact_off, &h_cursor,
                       &h_coords, &h_run_offs, &h_cigar, &h_moves};
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
    // Size of the variable-length results (compact strings / CIGAR runs) per pair seen in the last chunk of this
    // shape: the D2H copy of the next chunk is enqueued for that much (plus a margin) without a host round trip.
    int est_key[4] = {-1, -1, -1, -1};  // read_length, ref_length, mode, moves
    double est_per_pair = 0;

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
    bool moves = false;    // align results leave the device as CIGAR runs (packed entry points), not as strings
    bool zplane = false;   // SW align under the SSE/AVX pointer policy: the packed kernel stores a third plane
    bool affine = false;   // affine-gap variant: general kernel only, E boundary + 4-bit directions
    bool offsets = false;  // inputs are offset-addressed (packed entry points): per-pair lengths, no padding
    bool stage_reads = true, stage_refs = true, stage_offs = true;  // false: that input is page-locked, no staging copy
    size_t queue_words() const { return traceback_queue_words(read_length, ref_length); }
    // direction bytes per matrix row and pair: the general and the packed kernel keep separate
    // regions because one chunk can hold pairs of both kinds
    size_t gen_dir_row_bytes() const { return (size_t)segs * 2; }
    size_t dir_row_bytes() const { return gen_dir_row_bytes() + fast_dirs_bytes_per_row_per_slot(ref_length); }
    size_t per_pair_workspace() const {
        size_t b = (size_t)(read_chunks + ref_chunks) * 32 + (size_t)read_chunks * 8 + 2 * sizeof(PairMeta) + 40 + (size_t)rows_alloc * 6;
        if (align) {
            b += dir_row_bytes() * (rows_alloc + 4) + (size_t)ref_length * 2 + 8;
            if (zplane) b += fast_dirs_bytes_per_row_per_slot(ref_length) / 2 * (rows_alloc + 4);
            if (affine) b += (size_t)segs * 4 * rows_alloc;
            const size_t qw = traceback_queue_words(read_length, ref_length);
            if (traceback_wants_global_queue(read_length, ref_length)) b += qw * 4;
        }
        return b;
    }
    // worst-case bytes of one pair's compact strings: two strings of at most L moves + NUL
    size_t compact_bytes() const { return 2 * ((size_t)L + 1); }
    size_t per_pair_io() const {
        size_t b = (size_t)read_length + ref_length + 2 + 4 + (offsets ? 16 : 0);
        if (align) b += (moves ? 2 * (queue_words() + 1) * 4 + 24 : compact_bytes() + 4) + 2;
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
    if (alg == VA_OPT_SW || alg == VA_OPT_SW_AFFINE) return align ? MODE_SW_ALIGN : MODE_SW_SCORE;
    if (alg == VA_OPT_NW || alg == VA_OPT_NW_AFFINE) return align ? MODE_NW_ALIGN : MODE_NW_SCORE;
    return -1;
}
bool opt_is_affine(int opt) { return (opt & 0xF) == VA_OPT_SW_AFFINE || (opt & 0xF) == VA_OPT_NW_AFFINE; }
int opt_gap_open(int opt) { return -(int)(((unsigned)opt >> 8) & 0xFFFFu); }

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

// true when [p, p + bytes) is page-locked memory CUDA knows about (cudaHostAlloc / cudaHostRegister): the copy
// engines can read it in place, no staging copy needed
bool host_range_is_pinned(const void *p, size_t bytes) {
    if (!p || bytes == 0) return false;
    auto pinned = [](const void *q) {
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, q) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeHost;
    };
    return pinned(p) && pinned((const char *)p + bytes - 1);
}

}  // namespace

struct va_cuda_ctx {
    std::vector<Engine> engines;
    WorkerPool *pool = nullptr;
    int host_threads = 1;
    va_cuda_timings timings{};
    std::mutex call_mu;  // one call at a time per context (the reference's callers are single threaded)
};

namespace {

enum SlotKind { SLOT_RESIDENT = 0, SLOT_HOST = 1 };

int reserve_slot(ChunkSlot &s, const Shape &sh, int cap_pairs, SlotKind kind) {
    const bool pinned = kind == SLOT_HOST;
    const size_t slots = round_up((size_t)cap_pairs, 64);
    int rc;
    if (pinned) {  // the resident entry points read the caller's device buffers
        if ((rc = s.raw_reads.reserve((size_t)cap_pairs * sh.read_length + 16))) return rc;
        if ((rc = s.raw_refs.reserve((size_t)cap_pairs * sh.ref_length + 16))) return rc;
        if (sh.offsets) {
            if ((rc = s.read_off.reserve((slots + 1) * 8))) return rc;
            if ((rc = s.ref_off.reserve((slots + 1) * 8))) return rc;
        }
    }
    if ((rc = s.code_reads.reserve(slots * sh.read_chunks * 16 + 16))) return rc;
    if ((rc = s.code_refs.reserve(slots * sh.ref_chunks * 16 + 16))) return rc;
    if ((rc = s.row_idx.reserve(slots / 2 * sh.read_chunks * 16 + 16))) return rc;
    if ((rc = s.solo_list.reserve(slots * 4 + 64))) return rc;  // [0]: count, list from +16
    if ((rc = s.meta.reserve(slots * sizeof(PairMeta)))) return rc;
    if ((rc = s.pair_of.reserve(slots * 4))) return rc;
    if ((rc = s.prep_scratch.reserve(prep_scratch_bytes((int)slots, sh.read_length, sh.ref_length)))) return rc;
    // general region [rows][slots] + packed region [rows][duos] (kept apart: one chunk can hold both kinds)
    // (+16 rows: the tagged NW fill kernel stages the boundary column 16 rows at a time without a row test, va_nw.cu)
    if ((rc = s.boundary.reserve(slots * ((size_t)sh.rows_alloc + 16) * 6 + 512))) return rc;
    if ((rc = s.scores.reserve(slots * 2))) return rc;
    if ((rc = s.end_cell.reserve(slots * 4))) return rc;
    if (sh.affine && (rc = s.boundary_e.reserve(slots * sh.rows_alloc * 4 + 512))) return rc;
    if (sh.affine && sh.align && (rc = s.dirs4.reserve(slots * (size_t)sh.segs * 4 * sh.rows_alloc + 512))) return rc;
    if (sh.align) {
        if ((rc = s.dirs.reserve(slots * sh.dir_row_bytes() * (sh.rows_alloc + 4) + 512))) return rc;  // (+4: the intra-task layout pads rows to a multiple of 4)
        if (sh.zplane && (rc = s.zdirs.reserve(slots * (fast_dirs_bytes_per_row_per_slot(sh.ref_length) / 2) * (sh.rows_alloc + 4) + 512))) return rc;
        if ((rc = s.hrow.reserve(slots / 2 * (size_t)round_up((size_t)std::max(sh.ref_length, 1), 4) * 4 + 64))) return rc;
        // traceback move queue: shared memory unless the sequences are long
        const size_t qw = traceback_queue_words(sh.read_length, sh.ref_length);
        if (traceback_wants_global_queue(sh.read_length, sh.ref_length) && (rc = s.queue.reserve(slots * qw * 4 + 64))) return rc;
        if (pinned) {
            if ((rc = s.start.reserve(slots * 2))) return rc;
            if (sh.moves) {
                const size_t mv = (size_t)cap_pairs * (sh.queue_words() + 1) * 4 + 16;
                if ((rc = s.moves.reserve(mv))) return rc;
                if ((rc = s.cigar.reserve(mv))) return rc;
                if ((rc = s.coords.reserve(slots * 16))) return rc;
                if ((rc = s.run_count.reserve((slots + 1) * 4))) return rc;
                if ((rc = s.run_offs.reserve((slots + 1) * 4))) return rc;
                if ((rc = s.scan_tmp.reserve(cigar_compact_scratch_bytes((int)slots)))) return rc;
            } else {
                if ((rc = s.compact.reserve((size_t)cap_pairs * sh.compact_bytes() + 64))) return rc;
                if ((rc = s.compact_off.reserve(slots * 4))) return rc;
                if ((rc = s.cursor.reserve(64))) return rc;
            }
        }
    }
    if (pinned) {
        static const bool wc = [] { const char *v = getenv("VERSALIGN_CUDA_WC"); return v && atoi(v) != 0; }();
        if (sh.stage_reads && (rc = s.h_reads.reserve((size_t)cap_pairs * sh.read_length + 16, wc))) return rc;
        if (sh.stage_refs && (rc = s.h_refs.reserve((size_t)cap_pairs * sh.ref_length + 16, wc))) return rc;
        if (sh.offsets && sh.stage_offs) {
            if ((rc = s.h_read_off.reserve((slots + 1) * 8, wc))) return rc;
            if ((rc = s.h_ref_off.reserve((slots + 1) * 8, wc))) return rc;
        }
        if ((rc = s.h_scores.reserve(slots * 2))) return rc;
        if ((rc = s.h_end_cell.reserve(slots * 4))) return rc;
        if (sh.align) {
            if ((rc = s.h_start.reserve(slots * 2))) return rc;
            if (sh.moves) {
                if ((rc = s.h_cigar.reserve((size_t)cap_pairs * (sh.queue_words() + 1) * 4 + 16))) return rc;
                if ((rc = s.h_coords.reserve(slots * 16))) return rc;
                if ((rc = s.h_run_offs.reserve((slots + 1) * 4))) return rc;
            } else {
                if ((rc = s.h_compact.reserve((size_t)cap_pairs * sh.compact_bytes() + 64))) return rc;
                if ((rc = s.h_compact_off.reserve(slots * 4))) return rc;
                if ((rc = s.h_cursor.reserve(64))) return rc;
            }
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
    g.policy = 0;
    g.intra = 0;
    g.inband = 0;
    g.affine = 0;
    g.gap_open = 0;
}

// Where one chunk's device work reads and writes.
struct DeviceIO {
    const uint8_t *raw_reads = nullptr, *raw_refs = nullptr;
    const int64_t *read_off = nullptr, *ref_off = nullptr;  // offset-addressed input (device copies of the caller's offsets)
    int16_t *scores = nullptr, *end_cell = nullptr, *start = nullptr;
    uint8_t *aln_read = nullptr, *aln_ref = nullptr;  // fixed-stride strings (device-resident entry points)
    bool zero_prefix = false;                         // ... with the bytes before start[i] zeroed
    bool compact = false;                             // strings into ws.compact instead (host pipeline)
    bool moves = false;                               // CIGAR runs into ws.cigar + coordinates instead (packed entry points)
};

// Enqueue prep + fill (+ traceback) for n pairs.  Returns kernels launched (or < 0).
int enqueue_device_work(Engine &e, ChunkSlot &ws, const Shape &sh, int mode, int policy, const Scoring &sc, int gap_open, int n,
                        const DeviceIO &io, cudaStream_t stream, bool profile = false) {
#define ENQ_TRY(expr)                                                                                                  \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess)                                                                                         \
            return set_error(VA_ERR_DEVICE, "device %d: %s failed: %s", e.device, #expr, cudaGetErrorString(_e));       \
    } while (0)
    ChunkGeom g;
    fill_geom(g, sh, n);
    g.policy = sh.align ? policy : 0;
    g.affine = sh.affine ? 1 : 0;
    g.gap_open = gap_open;
    // VERSALIGN_CUDA_GENERAL_ONLY=1 keeps every pair on the 32-bit kernel (parity tests use it to
    // cover that kernel on inputs the packed kernels would otherwise take)
    static const bool general_only = [] { const char *v = getenv("VERSALIGN_CUDA_GENERAL_ONLY"); return v && atoi(v) != 0; }();
    // few, long pairs: the intra-task kernels (a CTA per pair-of-pairs, va_intra.cu); decided before prep: every
    // kernel of the chunk reads it
    static const bool no_intra = [] { const char *v = getenv("VERSALIGN_CUDA_NO_INTRA"); return v && atoi(v) != 0; }();
    if (!general_only && !sh.affine) {  // (the affine-gap variant runs on the general kernel)
        if (!no_intra && intra_preferred(mode, n, sh.read_length, sh.ref_length, e.sm_count) &&
            fast_scoring_ok(mode, policy, sc, sh.read_length, sh.ref_length, true)) {
            g.fast_tw = 16;
            g.intra = 1;
        } else if (fast_scoring_ok(mode, policy, sc, sh.read_length, sh.ref_length)) {
            g.fast_tw = fast_pick_tw(mode, sh.ref_length, g.policy);
            g.inband = fast_inband_ok(mode, policy, sc, sh.read_length, sh.ref_length) ? 1 : 0;
        }
    }
    const bool intra = g.intra != 0;
    ChunkBuffers b{};
    b.raw_reads = io.raw_reads;
    b.raw_refs = io.raw_refs;
    b.read_off = io.read_off;
    b.ref_off = io.ref_off;
    b.code_reads = (uint4 *)ws.code_reads.p;
    b.code_refs = (uint4 *)ws.code_refs.p;
    b.row_idx = (uint4 *)ws.row_idx.p;
    b.hrow = (uint32_t *)ws.hrow.p;
    b.solo_count = (int32_t *)ws.solo_list.p;
    b.solo_list = (int32_t *)ws.solo_list.p + 16;
    b.meta = (PairMeta *)ws.meta.p;
    b.pair_of = (int32_t *)ws.pair_of.p;
    b.boundary = (int32_t *)ws.boundary.p;
    b.boundary_e = (int32_t *)ws.boundary_e.p;
    b.dirs4 = (uint32_t *)ws.dirs4.p;
    b.fboundary = (uint32_t *)((char *)ws.boundary.p + round_up((size_t)g.slots * sh.rows_alloc * 4, 256));
    b.dirs = (uint16_t *)ws.dirs.p;
    b.fdirs = (uint4 *)((char *)ws.dirs.p + round_up((size_t)g.slots * sh.gen_dir_row_bytes() * sh.rows_alloc, 256));
    b.fdirs_z = sh.zplane ? (uint2 *)ws.zdirs.p : nullptr;
    b.scores = io.scores;
    b.end_cell = io.end_cell;
    b.aln_read = io.aln_read;
    b.aln_ref = io.aln_ref;
    b.start = io.start;
    if (io.moves) {
        b.moves_out = (uint32_t *)ws.moves.p;
        b.coords = (int32_t *)ws.coords.p;
        b.run_count = (uint32_t *)ws.run_count.p;
    }
    if (io.compact) {
        b.aln_compact = (uint8_t *)ws.compact.p;
        b.compact_off = (uint32_t *)ws.compact_off.p;
        b.compact_cursor = (unsigned long long *)ws.cursor.p;
    }
    b.cell_count = e.d_cells;
    int launches = 0;
    cudaEvent_t *pe = nullptr;
    if (profile) {
        if (e.prof_used + 4 > e.prof_events.size()) {
            for (int k = 0; k < 4; ++k) {
                cudaEvent_t ev;
                ENQ_TRY(cudaEventCreate(&ev));
                e.prof_events.push_back(ev);
            }
        }
        pe = &e.prof_events[e.prof_used];
        e.prof_used += 4;
        ENQ_TRY(cudaEventRecord(pe[0], stream));
    }
    // the inter-task kernels also take single slots (odd leftovers of the bucketing); the intra-task kernel
    // leaves those to the general kernel.  Decided before prep: every kernel of the chunk reads it.
    static const bool no_solo = [] { const char *v = getenv("VERSALIGN_CUDA_NO_SOLO"); return v && atoi(v) != 0; }();
    g.solo = (g.fast_tw && !intra && !no_solo) ? 1 : 0;
    launches += launch_prep(g, b, mode, policy, sc, ws.prep_scratch.p, ws.prep_scratch.cap, stream);
    if (pe) ENQ_TRY(cudaEventRecord(pe[1], stream));
    // The packed duo kernel, the solo kernel and the general kernel own disjoint slots: the general kernel
    // (usually with nothing to do) runs beside the packed ones on a side stream instead of after them.
    if (!ws.side) {
        ENQ_TRY(cudaStreamCreateWithFlags(&ws.side, cudaStreamNonBlocking));
        ENQ_TRY(cudaEventCreateWithFlags(&ws.ev_fork, cudaEventDisableTiming));
        ENQ_TRY(cudaEventCreateWithFlags(&ws.ev_join, cudaEventDisableTiming));
    }
    ENQ_TRY(cudaEventRecord(ws.ev_fork, stream));
    ENQ_TRY(cudaStreamWaitEvent(ws.side, ws.ev_fork, 0));
    launches += launch_fill_general(g, b, mode, policy, sc, ws.side);
    // bytes before start[i] are promised to be zero on the device-resident path: the copy engine clears the blocks
    // while the fill kernels (bound by the integer pipe) run -- 0.6 GB per million 150 bp pairs
    if (sh.align && io.zero_prefix && io.aln_read) {
        ENQ_TRY(cudaMemsetAsync(io.aln_read, 0, (size_t)n * sh.L, ws.side));
        ENQ_TRY(cudaMemsetAsync(io.aln_ref, 0, (size_t)n * sh.L, ws.side));
    }
    ENQ_TRY(cudaEventRecord(ws.ev_join, ws.side));
    if (intra)
        // (all four modes run in the Smith-Waterman form there: the SW tables and constants)
        launches += launch_fill_intra(g, b, mode, make_fast_consts(sh.align ? MODE_SW_ALIGN : MODE_SW_SCORE, sc), stream);
    else
        launches += launch_fill_fast(g, b, mode, sc, stream);
    ENQ_TRY(cudaStreamWaitEvent(stream, ws.ev_join, 0));
    if (pe) ENQ_TRY(cudaEventRecord(pe[2], stream));
    if (sh.align) {
        if (io.compact) ENQ_TRY(cudaMemsetAsync(ws.cursor.p, 0, 8, stream));
        if (io.moves) ENQ_TRY(cudaMemsetAsync((uint32_t *)ws.run_count.p + n, 0, 4, stream));  // the scan's sentinel entry
        launches += launch_traceback(g, b, mode, sc, (uint32_t *)ws.queue.p, stream);
        if (io.moves)
            launches += launch_cigar_compact(n, (int)sh.queue_words(), (const uint32_t *)ws.moves.p, (const uint32_t *)ws.run_count.p,
                                             (uint32_t *)ws.run_offs.p, (uint32_t *)ws.cigar.p, ws.cigar.cap / 4, ws.scan_tmp.p,
                                             ws.scan_tmp.cap, stream);
    }
    if (pe) ENQ_TRY(cudaEventRecord(pe[3], stream));
    ENQ_TRY(cudaGetLastError());
    return launches;
#undef ENQ_TRY
}

// CIGAR runs of one chunk of a packed align call
struct CigarPart {
    int64_t first = 0;
    int count = 0;
    size_t words = 0;
    bool global_offsets = false;  // cigar_off of these pairs already counts from run 0 of the batch
    std::unique_ptr<uint32_t[]> data;
};

// What the host-buffer entry points have in common.
struct HostCall {
    int mode = 0, policy = 0, gap_open = 0;
    Scoring sc{};
    Shape sh{};
    int n = 0;
    // inputs: either scattered or flat
    const char *const *reads_p = nullptr;
    const char *const *refs_p = nullptr;
    const char *reads_f = nullptr;
    const char *refs_f = nullptr;
    bool reads_pinned = false, refs_pinned = false, offs_pinned = false;  // flat inputs the copy engines can read in place
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
    int64_t *cigar_off = nullptr;                     // [n+1]; while the call runs: chunk-relative offsets, made global at the end
    std::vector<CigarPart> *cigar_parts = nullptr;  // one per chunk, any order
    int64_t *cigar_running = nullptr;                 // one device: chunks finish in pair order, so the runs before a chunk are known when it lands
    std::mutex *cigar_mu = nullptr;
    std::atomic<int> *cancel = nullptr;  // set by the first shard that fails: the others stop at their next chunk
};

struct ShardStats {
    double gather_s = 0, scatter_s = 0, kernel_ms = 0;
    int64_t h2d = 0, d2h = 0;
    int chunks = 0, launches = 0;
    unsigned long long cells = 0;
    int rc = VA_OK;
    std::string err;
};

// Inputs of pairs [first, first + count) -> pinned staging.  Returns false when there is nothing to stage
// (flat inputs in page-locked memory: the H2D copies read the caller's buffers in place).
void gather_chunk(va_cuda_ctx *ctx, const HostCall &c, ChunkSlot &s, int64_t first, int count) {
    char *hr = (char *)s.h_reads.p, *hf = (char *)s.h_refs.p;
    const int RL = c.sh.read_length, FL = c.sh.ref_length;
    if (c.read_off) {
        // offset-addressed sequences travel as they are: one block of bases per side + the offsets
        const int64_t r0 = c.read_off[first], r1 = c.read_off[first + count], f0 = c.ref_off[first], f1 = c.ref_off[first + count];
        if (!c.reads_pinned)
            ctx->pool->parallel_for(r1 - r0, 1 << 20, [&](int64_t b, int64_t e) { memcpy(hr + b, c.reads_f + r0 + b, (size_t)(e - b)); });
        if (!c.refs_pinned)
            ctx->pool->parallel_for(f1 - f0, 1 << 20, [&](int64_t b, int64_t e) { memcpy(hf + b, c.refs_f + f0 + b, (size_t)(e - b)); });
        if (!c.offs_pinned) {
            memcpy(s.h_read_off.p, c.read_off + first, (size_t)(count + 1) * 8);
            memcpy(s.h_ref_off.p, c.ref_off + first, (size_t)(count + 1) * 8);
        }
    } else if (c.reads_f) {
        if (!c.reads_pinned)
            ctx->pool->parallel_for((int64_t)count * RL, 1 << 20, [&](int64_t b, int64_t e) { memcpy(hr + b, c.reads_f + first * RL + b, (size_t)(e - b)); });
        if (!c.refs_pinned)
            ctx->pool->parallel_for((int64_t)count * FL, 1 << 20, [&](int64_t b, int64_t e) { memcpy(hf + b, c.refs_f + first * FL + b, (size_t)(e - b)); });
    } else {
        ctx->pool->parallel_for(count, 2048, [&](int64_t b, int64_t e) {
            // the blocks are scattered over the caller's heap: ask for the ones a few pairs ahead now
            constexpr int AHEAD = 8;
            static const bool stream_stores = [] { const char *v = getenv("VERSALIGN_CUDA_STREAM_STORES"); return !v || atoi(v) != 0; }();
            if (!stream_stores) {
                for (int64_t i = b; i < e; ++i) {
                    if (i + AHEAD < e) {
                        const char *pr = c.reads_p[first + i + AHEAD], *pf = c.refs_p[first + i + AHEAD];
                        for (int o = 0; o < RL; o += 64) __builtin_prefetch(pr + o, 0, 0);
                        for (int o = 0; o < FL; o += 64) __builtin_prefetch(pf + o, 0, 0);
                    }
                    memcpy(hr + i * RL, c.reads_p[first + i], RL);
                    memcpy(hf + i * FL, c.refs_p[first + i], FL);
                }
                return;
            }
            LineStreamer wr(hr + b * RL), wf(hf + b * FL);
            for (int64_t i = b; i < e; ++i) {
                if (i + AHEAD < e) {
                    const char *pr = c.reads_p[first + i + AHEAD], *pf = c.refs_p[first + i + AHEAD];
                    for (int o = 0; o < RL; o += 64) __builtin_prefetch(pr + o, 0, 0);
                    for (int o = 0; o < FL; o += 64) __builtin_prefetch(pf + o, 0, 0);
                }
                wr.append(c.reads_p[first + i], (size_t)RL);
                wf.append(c.refs_p[first + i], (size_t)FL);
            }
            wr.finish();
            wf.finish();
        });
    }
}

// Packed entry points, fallback: a chunk whose CIGARs do not fit the device-side compaction buffer (pairs with
// more runs than move slots) comes back as per-pair moves in walk order -- runs or the raw 2-bit queue
// (va_traceback.cu) -- and is replayed here.
void replay_moves_on_host(va_cuda_ctx *ctx, const HostCall &c, const uint32_t *mv, int64_t first, int count, CigarPart &cp) {
    const size_t qw = c.sh.queue_words() + 1;
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
    std::unique_ptr<int64_t[]> offs(new int64_t[(size_t)count + 1]);
    offs[0] = 0;
    ctx->pool->parallel_for(count, 8192, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) {
            const uint32_t head = mv[(size_t)i * qw];
            offs[(size_t)i + 1] = (head & 0x80000000u) ? replay(i, [](int, int) {}) : (int)head;
        }
    });
    for (int64_t i = 0; i < count; ++i) offs[(size_t)i + 1] += offs[(size_t)i];
    if (c.cigar_off)
        for (int64_t i = 0; i < count; ++i) c.cigar_off[first + i + 1] = offs[(size_t)i + 1];
    cp.words = (size_t)offs[(size_t)count];
    if (!c.cigar_parts) return;
    cp.data.reset(new uint32_t[cp.words + 1]);
    uint32_t *pout = cp.data.get();
    ctx->pool->parallel_for(count, 4096, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) {
            uint32_t *out = pout + offs[(size_t)i];
            replay(i, [&](int op, int len) { *out++ = ((uint32_t)len << 4) | (uint32_t)op; });
        }
    });
}

// Results of a packed align chunk -> the caller's arrays: everything was laid out on the device, the host only copies.
int scatter_packed(va_cuda_ctx *ctx, Engine &e, const HostCall &c, ChunkSlot &s, int64_t first, int count, ShardStats &st) {
    if (c.scores) memcpy(c.scores + first, s.h_scores.p, (size_t)count * sizeof(int16_t));
    if (c.end_cell) memcpy(c.end_cell + 2 * first, s.h_end_cell.p, (size_t)count * 2 * sizeof(int16_t));
    if (c.coords) {
        const char *src = (const char *)s.h_coords.p;
        ctx->pool->parallel_for((int64_t)count * 16, 1 << 20, [&](int64_t b, int64_t e2) { memcpy((char *)(c.coords + 4 * first) + b, src + b, (size_t)(e2 - b)); });
    }
    const uint32_t *ro = (const uint32_t *)s.h_run_offs.p;
    const size_t total = ro[count];
    CigarPart cp;
    cp.first = first;
    cp.count = count;
    // (the pinned block may be a little smaller than the device block: the two round their sizes differently)
    if (total > std::min(s.cigar.cap, s.h_cigar.cap) / 4) {
        // does not fit the compaction buffer (nothing was written past it): fetch the per-pair moves instead
        const size_t mv_bytes = (size_t)count * (c.sh.queue_words() + 1) * 4;
        int rc = s.h_moves.reserve(mv_bytes + 16);
        if (rc) return rc;
        cudaError_t err = cudaMemcpyAsync(s.h_moves.p, s.moves.p, mv_bytes, cudaMemcpyDeviceToHost, s.stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(s.stream);
        if (err != cudaSuccess) return set_error(VA_ERR_DEVICE, "device %d: fetching moves failed: %s", e.device, cudaGetErrorString(err));
        st.d2h += (int64_t)mv_bytes;
        replay_moves_on_host(ctx, c, (const uint32_t *)s.h_moves.p, first, count, cp);
    } else {
        if (total > s.sent) {  // more runs than the estimate the first copy was sized by: fetch the rest
            cudaError_t err = cudaMemcpyAsync((uint32_t *)s.h_cigar.p + s.sent, (const uint32_t *)s.cigar.p + s.sent, (total - s.sent) * 4,
                                              cudaMemcpyDeviceToHost, s.stream);
            if (err == cudaSuccess) err = cudaStreamSynchronize(s.stream);
            if (err != cudaSuccess) return set_error(VA_ERR_DEVICE, "device %d: fetching CIGAR runs failed: %s", e.device, cudaGetErrorString(err));
            st.d2h += (int64_t)(total - s.sent) * 4;
        }
        e.est_per_pair = (double)total / std::max(count, 1);
        const int64_t before = c.cigar_running ? *c.cigar_running : 0;
        cp.global_offsets = c.cigar_running != nullptr;
        if (c.cigar_off)
            ctx->pool->parallel_for(count, 1 << 14, [&](int64_t b, int64_t e2) {
                for (int64_t i = b; i < e2; ++i) c.cigar_off[first + i + 1] = before + (int64_t)ro[i + 1];
            });
        cp.words = total;
        if (c.cigar_parts) {
            cp.data.reset(new uint32_t[total + 1]);
            uint32_t *dst = cp.data.get();
            const uint32_t *src = (const uint32_t *)s.h_cigar.p;
            ctx->pool->parallel_for((int64_t)total, 1 << 18, [&](int64_t b, int64_t e2) { memcpy(dst + b, src + b, (size_t)(e2 - b) * 4); });
        }
    }
    if (c.cigar_running) *c.cigar_running += (int64_t)cp.words;
    if (c.cigar_parts) {
        std::lock_guard<std::mutex> lk(*c.cigar_mu);
        c.cigar_parts->push_back(std::move(cp));
    }
    return VA_OK;
}

// Results of a legacy align chunk: the two strings of pair i lie back to back in the compact block at off[i], each
// (moves + 1) bytes with its NUL; they go right-aligned into the caller's L-byte blocks (DefaultKernel.cpp:441-451).
int scatter_strings(va_cuda_ctx *ctx, Engine &e, const HostCall &c, ChunkSlot &s, int64_t first, int count, ShardStats &st) {
    const int L = c.sh.L;
    const int16_t *stt = (const int16_t *)s.h_start.p;
    const uint32_t *off = (const uint32_t *)s.h_compact_off.p;
    const char *hc = (const char *)s.h_compact.p;
    const size_t total = (size_t) * (const unsigned long long *)s.h_cursor.p;
    if (total > s.sent) {  // longer alignments than the estimate the first copy was sized by: fetch the rest
        cudaError_t err = cudaMemcpyAsync((char *)s.h_compact.p + s.sent, (const char *)s.compact.p + s.sent, total - s.sent,
                                          cudaMemcpyDeviceToHost, s.stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(s.stream);
        if (err != cudaSuccess) return set_error(VA_ERR_DEVICE, "device %d: fetching alignment strings failed: %s", e.device, cudaGetErrorString(err));
        st.d2h += (int64_t)(total - s.sent);
    }
    e.est_per_pair = (double)total / std::max(count, 1);
    if (c.start) memcpy(c.start + first, stt, (size_t)count * sizeof(int16_t));
    if (c.end_cell) memcpy(c.end_cell + 2 * first, s.h_end_cell.p, (size_t)count * 2 * sizeof(int16_t));
    // pair i: where the kept part of each string starts in the compact block, how many bytes (with the NUL), and
    // the index it goes to in the caller's block
    struct Piece {
        const char *a, *b;
        int s0, bytes;
    };
    auto piece = [&](int64_t i) {
        const int moves = L - 1 - (int)stt[i];
        int s0 = stt[i];
        if (s0 < 0) s0 = 0;  // a walk longer than the block (only when every move is a gap): its head is cut, like the fixed layout
        const int keep = L - 1 - s0;
        const char *a = hc + off[i] + (moves - keep);
        return Piece{a, a + moves + 1, s0, keep + 1};
    };
    if (c.out_read_f) {
        ctx->pool->parallel_for(count, 2048, [&](int64_t b, int64_t e2) {
            for (int64_t i = b; i < e2; ++i) {
                const Piece p = piece(i);
                char *da = c.out_read_f + (first + i) * L, *db = c.out_ref_f + (first + i) * L;
                memset(da, 0, (size_t)p.s0);
                memset(db, 0, (size_t)p.s0);
                memcpy(da + p.s0, p.a, (size_t)p.bytes);
                memcpy(db + p.s0, p.b, (size_t)p.bytes);
            }
        });
    } else if (c.alloc) {
        // Blocks are allocated a few pairs ahead of the copy that fills them, and the lines the strings will go to are
        // requested for writing at once: the allocator's work on pair i + AHEAD overlaps the memory latency of pair i
        static const int ahead_env = [] { const char *v = getenv("VERSALIGN_CUDA_ALLOC_AHEAD"); return v ? atoi(v) : 6; }();
        ctx->pool->parallel_for(count, 1024, [&](int64_t b, int64_t e2) {
            constexpr int MAX_AHEAD = 16;
            const int ahead = std::max(0, std::min(ahead_env, MAX_AHEAD - 1));
            char *ring_a[MAX_AHEAD], *ring_b[MAX_AHEAD];
            const size_t bytes = (size_t)(L > 0 ? L : 1);
            auto obtain = [&](int64_t i) {
                char *da = c.alloc(bytes, c.alloc_user);
                char *db = c.alloc(bytes, c.alloc_user);
                ring_a[i % MAX_AHEAD] = da;
                ring_b[i % MAX_AHEAD] = db;
                if (ahead > 0 && da && db) {
                    const int s0 = std::max(0, (int)stt[i]);
                    for (int o = s0; o < L; o += 64) {
                        __builtin_prefetch(da + o, 1, 3);
                        __builtin_prefetch(db + o, 1, 3);
                    }
                    __builtin_prefetch(da + L - 1, 1, 3);
                    __builtin_prefetch(db + L - 1, 1, 3);
                    __builtin_prefetch(hc + off[i], 0, 0);
                }
            };
            int64_t next = b;
            for (int64_t i = b; i < e2; ++i) {
                for (; next < e2 && next <= i + ahead; ++next) obtain(next);
                const Piece p = piece(i);
                char *da = ring_a[i % MAX_AHEAD], *db = ring_b[i % MAX_AHEAD];
                if (c.records) {
                    va_cuda_alignment_record *rec = reinterpret_cast<va_cuda_alignment_record *>(c.records + (size_t)(first + i) * c.record_stride);
                    rec->read = da;
                    rec->ref = db;
                    rec->read_start = rec->ref_start = stt[i];
                    rec->read_end = rec->ref_end = (int16_t)(L - 1);
                } else {
                    c.out_read_w[first + i] = da;
                    c.out_ref_w[first + i] = db;
                }
                if (!da || !db) {
                    c.alloc_failed->store(1);
                    continue;
                }
                memcpy(da + p.s0, p.a, (size_t)p.bytes);
                memcpy(db + p.s0, p.b, (size_t)p.bytes);
            }
        });
    } else {
        ctx->pool->parallel_for(count, 1024, [&](int64_t b, int64_t e2) {
            for (int64_t i = b; i < e2; ++i) {
                const Piece p = piece(i);
                memcpy(c.out_read_p[first + i] + p.s0, p.a, (size_t)p.bytes);
                memcpy(c.out_ref_p[first + i] + p.s0, p.b, (size_t)p.bytes);
            }
        });
    }
    return VA_OK;
}

int scatter_chunk(va_cuda_ctx *ctx, Engine &e, const HostCall &c, ChunkSlot &s, int64_t first, int count, ShardStats &st) {
    if (!c.sh.align) {
        memcpy(c.scores + first, s.h_scores.p, (size_t)count * sizeof(int16_t));
        return VA_OK;
    }
    return c.sh.moves ? scatter_packed(ctx, e, c, s, first, count, st) : scatter_strings(ctx, e, c, s, first, count, st);
}

bool long_pair_shape(int read_length, int ref_length);

// One device's share [lo, hi) of the batch, chunked through the ring.
void run_shard(va_cuda_ctx *ctx, Engine &e, const HostCall &c, int64_t lo, int64_t hi, int chunk_pairs, ShardStats &st) {
    // On any failure: remember the message (with the device), tell the sibling shards to stop, and leave nothing
    // in flight -- the ring's pinned buffers and the caller's outputs must not be written after the call returns.
    auto quiesce = [&] {
        for (auto &s : e.ring) {
            if (s.stream) cudaStreamSynchronize(s.stream);
            if (s.side) cudaStreamSynchronize(s.side);
            s.busy = false;
        }
        cudaGetLastError();
    };
    auto fail = [&](int rc) {
        st.rc = rc;
        st.err = g_last_error;
        if (c.cancel) c.cancel->store(1);
        quiesce();
    };
#define SHARD_TRY(expr)                                                                                                      \
    do {                                                                                                                     \
        cudaError_t _e = (expr);                                                                                             \
        if (_e != cudaSuccess)                                                                                               \
            return fail(set_error(_e == cudaErrorMemoryAllocation ? VA_ERR_MEMORY : VA_ERR_DEVICE, "device %d: %s failed: %s", \
                                  e.device, #expr, cudaGetErrorString(_e)));                                                 \
    } while (0)
    SHARD_TRY(cudaSetDevice(e.device));
    for (auto &s : e.ring) {
        int rc = reserve_slot(s, c.sh, chunk_pairs, SLOT_HOST);
        if (rc) return fail(rc);
        s.busy = false;
    }
    SHARD_TRY(cudaMemsetAsync(e.d_cells, 0, sizeof(unsigned long long), e.ring[0].stream));
    SHARD_TRY(cudaStreamSynchronize(e.ring[0].stream));

    const int RL = c.sh.read_length, FL = c.sh.ref_length;
    const bool strings = c.sh.align && !c.sh.moves, moves = c.sh.align && c.sh.moves;
    // per-pair size of the variable-length result, remembered from the last chunk of the same kind
    const int key[4] = {RL, FL, c.mode, c.sh.moves ? 1 : 0};
    if (memcmp(key, e.est_key, sizeof(key)) != 0) {
        memcpy(e.est_key, key, sizeof(key));
        e.est_per_pair = 0;  // unknown: the first chunk fetches its worst case
    }
    // VERSALIGN_CUDA_TRACE=1: host-side timeline of the shard (ms since the shard started)
    const bool trace = trace_on();
    const auto t_shard = Clock::now();
    auto mark = [&](const char *what, int64_t first) {
        if (trace) fprintf(stderr, "[va trace] dev %d %8.3f ms  %s %lld\n", e.device, seconds_since(t_shard) * 1e3, what, (long long)first);
    };
    auto drain = [&](ChunkSlot &s) -> int {
        if (!s.busy) return VA_OK;
        mark("wait", s.first);
        cudaError_t err = cudaEventSynchronize(s.ev_done);
        mark("done", s.first);
        if (err != cudaSuccess) return set_error(VA_ERR_DEVICE, "device %d: device work failed: %s", e.device, cudaGetErrorString(err));
        float ms = 0;
        if (cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1) == cudaSuccess) st.kernel_ms += ms;
        auto t0 = Clock::now();
        int rc = scatter_chunk(ctx, e, c, s, s.first, s.count, st);
        st.scatter_s += seconds_since(t0);
        mark("scattered", s.first);
        s.busy = false;
        if (rc == VA_OK && c.alloc_failed && c.alloc_failed->load()) rc = set_error(VA_ERR_MEMORY, "the caller's allocator returned NULL");
        return rc;
    };

    // Ramp: nothing leaves the device before the first chunk has been staged, copied, computed and copied back, and the
    // host has nothing to hand out until then.  A shard of several chunks therefore starts with a quarter and a half
    // chunk (legacy C2 call: first results after 2.3 instead of 4.3 ms of a 20 ms call).  Long pairs keep their
    // wave-sized chunks (a chunk is a handful of pairs there and the copies are nothing).
    static const bool ramp_allowed = [] { const char *v = getenv("VERSALIGN_CUDA_RAMP"); return !v || atoi(v) != 0; }();
    const bool ramp = ramp_allowed && hi - lo >= 3 * (int64_t)chunk_pairs && chunk_pairs >= 4096 && !long_pair_shape(RL, FL);
    int k = 0;
    int count = 0;
    for (int64_t first = lo; first < hi; first += count, ++k) {
        if (c.cancel && c.cancel->load()) {  // another device's shard failed
            quiesce();
            return;
        }
        int64_t want = chunk_pairs;
        if (ramp && k < 2) want = (int64_t)round_up((size_t)(chunk_pairs >> (2 - k)), 64);
        count = (int)std::min<int64_t>(want, hi - first);
        ChunkSlot &s = e.ring[k % kRing];
        int rc = drain(s);  // the slot's previous chunk must be out before its pinned buffers are reused
        if (rc) return fail(rc);
        auto t0 = Clock::now();
        gather_chunk(ctx, c, s, first, count);
        st.gather_s += seconds_since(t0);
        mark("gathered", first);

        DeviceIO io;
        io.raw_reads = (const uint8_t *)s.raw_reads.p;
        io.raw_refs = (const uint8_t *)s.raw_refs.p;
        if (c.read_off) {
            const int64_t r0 = c.read_off[first], r1 = c.read_off[first + count], f0 = c.ref_off[first], f1 = c.ref_off[first + count];
            SHARD_TRY(cudaMemcpyAsync(s.raw_reads.p, c.reads_pinned ? (const void *)(c.reads_f + r0) : s.h_reads.p, (size_t)(r1 - r0), cudaMemcpyHostToDevice, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.raw_refs.p, c.refs_pinned ? (const void *)(c.refs_f + f0) : s.h_refs.p, (size_t)(f1 - f0), cudaMemcpyHostToDevice, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.read_off.p, c.offs_pinned ? (const void *)(c.read_off + first) : s.h_read_off.p, (size_t)(count + 1) * 8, cudaMemcpyHostToDevice, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.ref_off.p, c.offs_pinned ? (const void *)(c.ref_off + first) : s.h_ref_off.p, (size_t)(count + 1) * 8, cudaMemcpyHostToDevice, s.stream));
            st.h2d += (r1 - r0) + (f1 - f0) + (int64_t)(count + 1) * 16;
            io.read_off = (const int64_t *)s.read_off.p;
            io.ref_off = (const int64_t *)s.ref_off.p;
        } else {
            const bool rp = c.reads_f && c.reads_pinned, fp = c.refs_f && c.refs_pinned;
            SHARD_TRY(cudaMemcpyAsync(s.raw_reads.p, rp ? (const void *)(c.reads_f + first * RL) : s.h_reads.p, (size_t)count * RL, cudaMemcpyHostToDevice, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.raw_refs.p, fp ? (const void *)(c.refs_f + first * FL) : s.h_refs.p, (size_t)count * FL, cudaMemcpyHostToDevice, s.stream));
            st.h2d += (int64_t)count * (RL + FL);
        }
        SHARD_TRY(cudaEventRecord(s.ev_k0, s.stream));
        io.scores = (int16_t *)s.scores.p;
        io.end_cell = (int16_t *)s.end_cell.p;
        io.start = (int16_t *)s.start.p;
        io.compact = strings;
        io.moves = moves;
        int launches = enqueue_device_work(e, s, c.sh, c.mode, c.policy, c.sc, c.gap_open, count, io, s.stream);
        if (launches < 0) return fail(launches);
        st.launches += launches;
        SHARD_TRY(cudaEventRecord(s.ev_k1, s.stream));
        if (!c.sh.align) {
            SHARD_TRY(cudaMemcpyAsync(s.h_scores.p, s.scores.p, (size_t)count * 2, cudaMemcpyDeviceToHost, s.stream));
            st.d2h += (int64_t)count * 2;
        } else if (moves) {
            // scores, coordinates, run offsets; the runs themselves for as many as the last chunk had per pair (+ margin)
            const size_t cap_words = std::min(s.cigar.cap, s.h_cigar.cap) / 4;
            const size_t want = e.est_per_pair > 0 ? (size_t)(e.est_per_pair * 1.05 * count) + 1024 : cap_words;
            s.sent = std::min(cap_words, std::min(want, (size_t)count * (c.sh.queue_words() + 1)));
            SHARD_TRY(cudaMemcpyAsync(s.h_run_offs.p, s.run_offs.p, (size_t)(count + 1) * 4, cudaMemcpyDeviceToHost, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.h_coords.p, s.coords.p, (size_t)count * 16, cudaMemcpyDeviceToHost, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.h_end_cell.p, s.end_cell.p, (size_t)count * 4, cudaMemcpyDeviceToHost, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.h_scores.p, s.scores.p, (size_t)count * 2, cudaMemcpyDeviceToHost, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.h_cigar.p, s.cigar.p, s.sent * 4, cudaMemcpyDeviceToHost, s.stream));
            st.d2h += (int64_t)count * 26 + 4 + (int64_t)s.sent * 4;
        } else {
            const size_t cap_bytes = (size_t)count * c.sh.compact_bytes();
            const size_t want = e.est_per_pair > 0 ? (size_t)(e.est_per_pair * 1.03 * count) + 4096 : cap_bytes;
            s.sent = std::min(cap_bytes, want);
            SHARD_TRY(cudaMemcpyAsync(s.h_cursor.p, s.cursor.p, 8, cudaMemcpyDeviceToHost, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.h_start.p, s.start.p, (size_t)count * 2, cudaMemcpyDeviceToHost, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.h_compact_off.p, s.compact_off.p, (size_t)count * 4, cudaMemcpyDeviceToHost, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.h_end_cell.p, s.end_cell.p, (size_t)count * 4, cudaMemcpyDeviceToHost, s.stream));
            SHARD_TRY(cudaMemcpyAsync(s.h_compact.p, s.compact.p, s.sent, cudaMemcpyDeviceToHost, s.stream));
            st.d2h += (int64_t)count * 10 + 8 + (int64_t)s.sent;
        }
        SHARD_TRY(cudaEventRecord(s.ev_done, s.stream));
        mark("enqueued", first);
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
    SHARD_TRY(cudaMemcpy(&cells, e.d_cells, sizeof(cells), cudaMemcpyDeviceToHost));
    st.cells = cells;
    SHARD_TRY(cudaGetLastError());
#undef SHARD_TRY
}

// Long-pair align chunks (intra-task kernels, va_intra.cu): a CTA per pair-of-pairs with 4, 8 or 16 warps, 16 warps per
// SM -- so the device holds 4, 2 or 1 x sm_count duos at a time and a chunk should be exactly one such wave.  And the
// direction words one launch touches should stay below ~24 GB: at 35 GB (1184 pairs of 10 kbp x 12 kbp) the fill ran 20 %
// and the traceback 3x slower than at 18 GB -- the accesses are spread over more pages than the TLBs reach.
int64_t intra_align_chunk_pairs(const Engine &e, const Shape &sh) {
    const size_t per_duo = (size_t)sh.ref_chunks * 8 * (((size_t)sh.rows_alloc + 3) & ~(size_t)3);
    const size_t limit = (size_t)24 << 30;
    for (int waves : {4, 2, 1}) {
        const int64_t duos = (int64_t)e.sm_count * waves;
        if ((size_t)duos * per_duo <= limit || waves == 1) return 2 * duos;
    }
    return 2 * (int64_t)e.sm_count;
}

bool long_pair_shape(int read_length, int ref_length) { return read_length >= 256 && ref_length >= 1024; }

int pick_chunk_pairs(const Engine &e, const Shape &sh, int64_t shard_pairs) {
    // device workspace budget per ring slot: a quarter of a third of the card, at most 6 GiB
    // (long pairs: tens of MB of directions per pair and next to nothing to copy -- fewer, larger chunks keep the
    // intra-task kernels' CTAs-per-duo grid full)
    const size_t per_pair = sh.per_pair_workspace() + sh.per_pair_io();
    const size_t budget = per_pair > ((size_t)4 << 20) ? std::min<size_t>(e.total_mem / 6, (size_t)24 << 30) : std::min<size_t>(e.total_mem / 12, (size_t)6 << 30);
    int64_t cap = (int64_t)std::max<size_t>(64, budget / std::max<size_t>(per_pair, 1));
    // pinned staging per slot at most ~512 MiB
    cap = std::min<int64_t>(cap, std::max<int64_t>(64, ((int64_t)512 << 20) / (int64_t)std::max<size_t>(sh.per_pair_io(), 1)));
    // aim for >= 8 chunks per shard so the pipeline overlaps, but keep chunks >= min_chunk pairs (a chunk
    // costs a fixed ~0.3 ms of launches, copies and hand-offs)
    static const int64_t min_chunk = [] { const char *v = getenv("VERSALIGN_CUDA_MIN_CHUNK"); return v && atoll(v) > 0 ? atoll(v) : 65536LL; }();
    int64_t want = std::max<int64_t>((shard_pairs + 7) / 8, min_chunk);
    // A chunk's fill kernel should be whole waves of the device: the packed kernels run one thread per pair-of-pairs, 512
    // threads per SM, so one wave is sm_count * 1024 pairs (151 552 on a B200).  A 125 k-pair chunk is 0.82 of a wave
    // and pays for a whole one: 8 -> 7 chunks per million pairs took the C2 call from 23.9 to 20.9 ms (legacy boundary)
    // and from 13.1 to 10.7 ms (packed boundary).
    const int64_t wave = (int64_t)e.sm_count * 1024;
    if (want * 2 >= wave) want = std::max<int64_t>(1, (want + wave / 2) / wave) * wave;
    want = std::min<int64_t>(want, cap);
    // (only when a chunk of the usual size would go to the intra-task kernels anyway)
    if (sh.align && long_pair_shape(sh.read_length, sh.ref_length) &&
        intra_preferred(MODE_SW_ALIGN, (int)std::min<int64_t>(want, shard_pairs), sh.read_length, sh.ref_length, e.sm_count))
        want = std::min<int64_t>(cap, intra_align_chunk_pairs(e, sh));
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
        std::atomic<int> cancel{0};
        c.cancel = &cancel;
        // flat inputs in page-locked memory are read in place by the copy engines
        if (c.reads_f && !c.reads_p) {
            const size_t rbytes = c.read_off ? (size_t)(c.read_off[c.n] - c.read_off[0]) : (size_t)c.n * c.sh.read_length;
            const size_t fbytes = c.ref_off ? (size_t)(c.ref_off[c.n] - c.ref_off[0]) : (size_t)c.n * c.sh.ref_length;
            c.reads_pinned = host_range_is_pinned(c.reads_f + (c.read_off ? c.read_off[0] : 0), rbytes);
            c.refs_pinned = host_range_is_pinned(c.refs_f + (c.ref_off ? c.ref_off[0] : 0), fbytes);
            c.offs_pinned = c.read_off && host_range_is_pinned(c.read_off, (size_t)(c.n + 1) * 8) && host_range_is_pinned(c.ref_off, (size_t)(c.n + 1) * 8);
            c.sh.stage_reads = !c.reads_pinned;
            c.sh.stage_refs = !c.refs_pinned;
            c.sh.stage_offs = !c.offs_pinned;
        }
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
        call_mark("shards start");
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
        call_mark("shards done");
        c.cancel = nullptr;
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
    call_begin();  // (trace marks count from here)
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
    c.sh.affine = opt_is_affine(opt);
    c.gap_open = c.sh.affine ? opt_gap_open(opt) : 0;
    if (c.sh.affine && align && policy != VA_POLICY_DEFAULT_OCL) return set_error(VA_ERR_ARG, "the affine-gap variant has one pointer rule (policy 0)");
    c.sh.zplane = align && policy == VA_POLICY_SIMD && c.mode == MODE_SW_ALIGN && !c.sh.affine;
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
static int packed_lengths(va_cuda_ctx *ctx, int n, const int64_t *read_off, const int64_t *ref_off, int *read_length, int *ref_length) {
    if (!ctx) return set_error(VA_ERR_ARG, "context is null");
    std::atomic<int64_t> rl{0}, fl{0};
    std::atomic<int> bad{-1};
    ctx->pool->parallel_for(n, 1 << 16, [&](int64_t b, int64_t e) {
        int64_t r = 0, f = 0;
        for (int64_t i = b; i < e; ++i) {
            const int64_t a = read_off[i + 1] - read_off[i], c = ref_off[i + 1] - ref_off[i];
            if (a < 0 || c < 0) bad.store((int)i);
            r = std::max(r, a);
            f = std::max(f, c);
        }
        int64_t cur = rl.load();
        while (r > cur && !rl.compare_exchange_weak(cur, r)) {
        }
        cur = fl.load();
        while (f > cur && !fl.compare_exchange_weak(cur, f)) {
        }
    });
    if (bad.load() >= 0) return set_error(VA_ERR_ARG, "offsets of pair %d decrease", bad.load());
    if (rl.load() > 32000 || fl.load() > 32000) return set_error(VA_ERR_RANGE, "a sequence is longer than 32000 bases");
    *read_length = (int)rl.load();
    *ref_length = (int)fl.load();
    return VA_OK;
}

int va_cuda_score_packed(va_cuda_ctx *ctx, int opt, const va_cuda_scoring *sc, int n, const char *reads,
                         const int64_t *read_off, const char *refs, const int64_t *ref_off, int16_t *scores) {
    if (n < 0) return set_error(VA_ERR_ARG, "n < 0");
    if (n > 0 && (!reads || !refs || !read_off || !ref_off || !scores)) return set_error(VA_ERR_ARG, "null buffer");
    int rl = 0, fl = 0;
    int rc = packed_lengths(ctx, n, read_off, ref_off, &rl, &fl);
    if (rc) return rc;
    HostCall c;
    bool noop;
    rc = prepare_call(ctx, c, opt, false, 0, sc, n, rl, fl, &noop);
    if (rc || noop) return rc;
    c.sh.offsets = true;
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
    int rc = packed_lengths(ctx, n, read_off, ref_off, &rl, &fl);
    if (rc) return rc;
    HostCall c;
    bool noop;
    rc = prepare_call(ctx, c, opt, true, policy, sc, n, rl, fl, &noop);
    if (rc || noop) return rc;
    c.sh.moves = true;
    c.sh.offsets = true;
    c.reads_f = reads;
    c.refs_f = refs;
    c.read_off = read_off;
    c.ref_off = ref_off;
    c.scores = scores;
    c.coords = coords;
    c.cigar_off = cigar_off;
    std::vector<CigarPart> parts;
    std::mutex mu;
    int64_t running = 0;
    if (cigar || cigar_off) {
        c.cigar_parts = &parts;
        c.cigar_mu = &mu;
        if (ctx->engines.size() == 1) c.cigar_running = &running;
    }
    if (cigar_off) cigar_off[0] = 0;
    rc = run_host_call(ctx, c);
    if (rc) return rc;
    // every chunk left its pairs' offsets relative to its own first run: line the chunks up in pair order
    call_mark("run_host_call returned");
    std::sort(parts.begin(), parts.end(), [](const CigarPart &a, const CigarPart &b) { return a.first < b.first; });
    std::vector<int64_t> base(parts.size() + 1, 0);
    for (size_t k = 0; k < parts.size(); ++k) base[k + 1] = base[k] + (int64_t)parts[k].words;
    if (cigar_off) {
        for (size_t k = 1; k < parts.size(); ++k) {  // chunk 0 starts at run 0 already
            if (parts[k].global_offsets) continue;
            int64_t *o = cigar_off + parts[k].first + 1;
            const int64_t add = base[k];
            ctx->pool->parallel_for(parts[k].count, 1 << 16, [&](int64_t b, int64_t e) {
                for (int64_t i = b; i < e; ++i) o[i] += add;
            });
        }
    }
    call_mark("offsets lined up");
    if (cigar) {
        const int64_t total = base[parts.size()];
        uint32_t *out = (uint32_t *)alloc((size_t)std::max<int64_t>(total, 1) * sizeof(uint32_t), user);
        if (!out) return set_error(VA_ERR_MEMORY, "the caller's allocator returned NULL");
        // one pass over the whole block: piece [b, e) of the output may span several parts
        ctx->pool->parallel_for(total, 1 << 17, [&](int64_t b, int64_t e) {
            size_t k = (size_t)(std::upper_bound(base.begin(), base.end(), b) - base.begin()) - 1;
            while (b < e) {
                const int64_t stop = std::min<int64_t>(e, base[k + 1]);
                if (stop > b) memcpy(out + b, parts[k].data.get() + (b - base[k]), (size_t)(stop - b) * 4);
                b = std::max(b, stop);
                ++k;
            }
        });
        *cigar = out;
    }
    call_mark("runs copied");
    return VA_OK;
}

int va_cuda_max_resident_pairs(va_cuda_ctx *ctx, int align, int read_length, int ref_length, int64_t *max_n) {
    if (!ctx || !max_n) return set_error(VA_ERR_ARG, "null argument");
    const Shape sh = make_shape(read_length, ref_length, align != 0);
    const Engine &e = ctx->engines[0];
    // (an align sub-chunk's direction region stays below 2^32 eight-byte words: the traceback walk's offsets are 32-bit)
    // (long pairs go through the intra-task layout and the warp-per-pair traceback, which index with 64 bits)
    const bool long_pairs = long_pair_shape(read_length, ref_length);
    const size_t budget = std::min<size_t>(e.total_mem / 2, align && !long_pairs ? (size_t)30 << 30 : (size_t)80 << 30);
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
    // the resident workspace, the side stream and the profiling events are shared by all calls on the context
    std::lock_guard<std::mutex> lk(ctx->call_mu);
    Engine &e = ctx->engines[0];
    CUDA_TRY(cudaSetDevice(e.device));
    int64_t max_n = 0;
    va_cuda_max_resident_pairs(ctx, align, read_length, ref_length, &max_n);
    int64_t sub64 = std::max<int64_t>(64, max_n / 64 * 64);
    if (align && long_pair_shape(read_length, ref_length) && intra_preferred(c.mode, (int)std::min<int64_t>(n, sub64), read_length, ref_length, e.sm_count))
        sub64 = std::min<int64_t>(sub64, intra_align_chunk_pairs(e, c.sh));
    const int sub = (int)std::min<int64_t>(n, sub64);
    rc = reserve_slot(e.resident, c.sh, sub, SLOT_RESIDENT);
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
        DeviceIO io;
        io.raw_reads = (const uint8_t *)d_reads + first * read_length;
        io.raw_refs = (const uint8_t *)d_refs + first * ref_length;
        io.scores = d_scores ? (int16_t *)d_scores + first : (int16_t *)ws.scores.p + first;
        io.end_cell = d_end_cell ? (int16_t *)d_end_cell + 2 * first : (int16_t *)ws.end_cell.p + (align ? 2 * first : 0);
        if (align) {
            io.aln_read = (uint8_t *)d_aln_read + first * L;
            io.aln_ref = (uint8_t *)d_aln_ref + first * L;
            io.start = (int16_t *)d_start + first;
            io.zero_prefix = true;
        }
        int k = enqueue_device_work(e, ws, c.sh, c.mode, c.policy, c.sc, c.gap_open, count, io, st, e.profiling);
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
    std::lock_guard<std::mutex> lk(ctx->call_mu);
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
    std::lock_guard<std::mutex> lk(ctx->call_mu);
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
