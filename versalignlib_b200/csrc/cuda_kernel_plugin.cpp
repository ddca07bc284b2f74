// cuda_kernel_plugin.cpp -- "CUDAKernel": one more AlignmentKernel behind the reference's
// plug-in boundary, so the reference driver loads libCUDAKernel.so exactly like its Default,
// SSE, AVX and OpenCL libraries (add { "CUDA", libPathDir + "libCUDAKernel" + libSuffix } to
// kernel_map, src/impl/main.cpp:61-64).
//
//   reference                                              here
//   DefaultKernel.h:66-81   ctor: six required keys, throws a C string when one is missing   CUDAKernel::CUDAKernel
//   DefaultKernel.cpp:52-81 score_alignments: opt&0xF dispatch, log, run                     CUDAKernel::score_alignments
//   DefaultKernel.cpp:21-50 compute_alignments, :441-451 new char[alnLength] + four shorts   CUDAKernel::compute_alignments
//   DefaultKernel_dllexport.cpp:18-42  the four extern "C" symbols + _parameters/_logger     bottom of this file
//
// All device work goes through the flat C ABI (versalign_cuda.h); there is no CPU path.
#include <cstddef>
#include <malloc.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "versalign_cuda.h"
#include "versalign_plugin_abi.h"

#define KERNEL "CUDA"

AlignmentParameters *_parameters = 0;
AlignmentLogger *_logger = 0;

namespace {

// Optional keys are probed with has_key() first: the reference driver's param_int() throws on
// unknown keys (CustomParameters.h:25-27).  Environment variables are the fallback so the
// stock driver can steer the plug-in without code changes.
int optional_param(const char *key, const char *env, int fallback) {
    if (_parameters && Parameters.has_key(key)) return Parameters.param_int(key);
    if (const char *v = getenv(env)) return atoi(v);
    return fallback;
}

void log_info(const std::string &msg) {
    if (_logger) Logger.log(0, KERNEL, msg.c_str());
}

[[noreturn]] void fatal(const std::string &msg) {
    // reference convention: log at level 3, then throw (OpenCLKernel.cpp:647-655); we throw a
    // C string like the constructors do so hosts can catch (const char*)
    static thread_local std::string keep;
    keep = msg;
    if (_logger) Logger.log(3, KERNEL, keep.c_str());
    throw keep.c_str();
}

// The library is never dlclose()d by the reference driver (main.cpp:217-225 re-dlopens to find
// delete_alignment_kernel), and it spawns a fresh kernel per timing run (main.cpp:261-265), so
// the CUDA context (streams, pinned staging, device workspace) is shared by all instances.
// One context per (cuda_devices, cuda_device_first) combination, created on first use and kept for
// the life of the process: instances spawned earlier with another combination keep a valid ctx_.
struct SharedContext {
    std::mutex mu;
    struct Entry {
        int n_devices, first;
        va_cuda_ctx *ctx;
    };
    std::vector<Entry> entries;
    va_cuda_ctx *last = nullptr;  // context of the most recent call (va_cuda_plugin_timings)
};
SharedContext g_shared;

va_cuda_ctx *acquire_context(int n_devices, int first) {
    std::lock_guard<std::mutex> lk(g_shared.mu);
    for (const auto &e : g_shared.entries)
        if (e.n_devices == n_devices && e.first == first) return g_shared.last = e.ctx;
    int visible = 0;
    if (va_cuda_device_count(&visible) != VA_OK) fatal(std::string("Cannot instantiate Kernel. ") + va_cuda_last_error());
    std::vector<int> devs;
    int f = first;
    if (f < 0 || f >= visible) f = 0;
    const int use = (n_devices <= 0 || f + n_devices > visible) ? visible - f : n_devices;
    for (int d = 0; d < use; ++d) devs.push_back(f + d);
    va_cuda_ctx *ctx = nullptr;
    if (va_cuda_create(&ctx, devs.data(), (int)devs.size(), 0) != VA_OK)
        fatal(std::string("Cannot instantiate Kernel. ") + va_cuda_last_error());
    g_shared.entries.push_back({n_devices, first, ctx});
    return g_shared.last = ctx;
}

// compute_alignments hands back two new char[] blocks per pair (the interface demands it:
// AlignmentKernel.h:20-23 delete[]s them one by one), and the caller frees them after every call.  With
// glibc's defaults every freed heap top goes back to the kernel at once and the next call page-faults
// hundreds of MB in again -- on the staging threads, serialised on the process's mmap lock: with 4
// threads that is 4-5x the cost of the allocation itself (tools/malloc_probe.cpp).  So the plug-in asks
// glibc once to keep freed pages in its arenas.  VERSALIGN_CUDA_MALLOC_TUNE=0 leaves the allocator alone.
void tune_host_malloc_once() {
    static std::once_flag once;
    std::call_once(once, [] {
        const char *v = getenv("VERSALIGN_CUDA_MALLOC_TUNE");
        if (v && atoi(v) == 0) return;
        mallopt(M_TRIM_THRESHOLD, 1 << 30);
        mallopt(M_TOP_PAD, 64 << 20);
        mallopt(M_MMAP_THRESHOLD, 1 << 30);
    });
}

class CUDAKernel : public AlignmentKernel {
public:
    CUDAKernel() {
        bool missing = _parameters == 0;
        auto need = [&](const char *key) -> int {
            if (missing || !Parameters.has_key(key)) {
                missing = true;
                return 0;
            }
            return Parameters.param_int(key);
        };
        // values are narrowed to short like the reference's members (DefaultKernel.h:139-142)
        scoring_.match = (short)need("score_match");
        scoring_.mismatch = (short)need("score_mismatch");
        scoring_.gap_read = (short)need("score_gap_read");
        scoring_.gap_ref = (short)need("score_gap_ref");
        read_length_ = need("read_length");
        ref_length_ = need("ref_length");
        if (missing) throw "Cannot instantiate Kernel. Lacking parameters";
        policy_ = optional_param("cuda_traceback_policy", "VERSALIGN_CUDA_POLICY", VA_POLICY_DEFAULT_OCL);
        if (policy_ != VA_POLICY_DEFAULT_OCL && policy_ != VA_POLICY_SIMD) fatal("cuda_traceback_policy must be 0 (Default/OpenCL) or 1 (SSE/AVX)");
        gap_open_ = optional_param("score_gap_open", "VERSALIGN_CUDA_GAP_OPEN", 0);
        if (gap_open_ > 0 || gap_open_ < -32767) fatal("score_gap_open must be in [-32767, 0]");
        const int n_devices = optional_param("cuda_devices", "VERSALIGN_CUDA_DEVICES", 0);
        // one process per GPU (torchrun): rank r passes cuda_device_first = r, cuda_devices = 1
        const int first = optional_param("cuda_device_first", "VERSALIGN_CUDA_DEVICE_FIRST", 0);
        ctx_ = acquire_context(n_devices, first);
        tune_host_malloc_once();
        log_info("Successfully instantiated CUDA Kernel.");
    }

    ~CUDAKernel() override {}

    void score_alignments(int const &opt, int const &aln_number, char const *const *const reads,
                          char const *const *const refs, short *const scores) override {
        const int alg = opt & 0xF;
        if (alg < VA_OPT_SW || alg > VA_OPT_NW_AFFINE) return;  // unsupported mode: touch nothing
        apply_threads();
        log_info("Running CUDAKernel score.");
        int rc = va_cuda_score_ptrs(ctx_, with_gap_open(opt), &scoring_, aln_number, reads, read_length_, refs, ref_length_, scores);
        if (rc != VA_OK) fatal(std::string("score_alignments failed: ") + va_cuda_last_error());
    }

    void compute_alignments(int const &opt, int const &aln_number, char const *const *const reads,
                            char const *const *const refs, Alignment *const alignments) override {
        const int alg = opt & 0xF;
        if (alg < VA_OPT_SW || alg > VA_OPT_NW_AFFINE) return;
        apply_threads();
        log_info("Running CUDAKernel align.");
        const int n = aln_number;
        if (n <= 0) return;
        // Result blocks must be individually delete[]-able (Alignment::~Alignment), so each is a
        // plain array-new block; the C ABI calls this allocator from its staging threads while
        // results stream back from the device, and fills the Alignment records there too (previous
        // contents are neither freed nor reused: reference semantics).
        static_assert(sizeof(Alignment) >= sizeof(va_cuda_alignment_record), "Alignment layout");
        static_assert(offsetof(Alignment, read) == offsetof(va_cuda_alignment_record, read) &&
                          offsetof(Alignment, ref) == offsetof(va_cuda_alignment_record, ref) &&
                          offsetof(Alignment, readStart) == offsetof(va_cuda_alignment_record, read_start) &&
                          offsetof(Alignment, readEnd) == offsetof(va_cuda_alignment_record, read_end) &&
                          offsetof(Alignment, refStart) == offsetof(va_cuda_alignment_record, ref_start) &&
                          offsetof(Alignment, refEnd) == offsetof(va_cuda_alignment_record, ref_end),
                      "va_cuda_alignment_record restates struct Alignment (AlignmentKernel.h:12-24)");
        int rc = va_cuda_align_records(ctx_, with_gap_open(opt), alg >= VA_OPT_SW_AFFINE ? VA_POLICY_DEFAULT_OCL : policy_, &scoring_, n, reads, read_length_, refs, ref_length_,
                                       [](size_t bytes, void *) -> char * { return new (std::nothrow) char[bytes]; }, nullptr,
                                       alignments, sizeof(Alignment), nullptr);
        if (rc != VA_OK) fatal(std::string("compute_alignments failed: ") + va_cuda_last_error());
    }

private:
    // opt values 2 / 3 (beyond the reference's 0 / 1): the affine-gap variants; their gap-open score is the optional
    // key "score_gap_open" (env VERSALIGN_CUDA_GAP_OPEN), read when the kernel is spawned
    int with_gap_open(int opt) const { return (opt & 0xF) >= VA_OPT_SW_AFFINE ? ((opt & 0xF) | VA_OPT_GAP_OPEN(gap_open_)) : opt; }

    // num_threads is read on every call like the reference (DefaultKernel.cpp:45); here it sizes
    // the host staging pool.
    void apply_threads() {
        int t = 0;
        if (_parameters && Parameters.has_key("num_threads")) t = Parameters.param_int("num_threads");
        const int env = optional_param("cuda_host_threads", "VERSALIGN_CUDA_HOST_THREADS", 0);
        if (env > 0) t = env;
        va_cuda_set_host_threads(ctx_, t);
        std::lock_guard<std::mutex> lk(g_shared.mu);
        g_shared.last = ctx_;
    }

    va_cuda_scoring scoring_{};
    int read_length_ = 0, ref_length_ = 0;
    int policy_ = VA_POLICY_DEFAULT_OCL;
    int gap_open_ = 0;
    va_cuda_ctx *ctx_ = nullptr;
};

}  // namespace

extern "C" int va_cuda_plugin_timings(va_cuda_timings *out) {
    std::lock_guard<std::mutex> lk(g_shared.mu);
    if (!g_shared.last || !out) return VA_ERR_ARG;
    return va_cuda_get_timings(g_shared.last, out);
}

extern "C" AlignmentKernel *spawn_alignment_kernel() { return new CUDAKernel(); }

extern "C" void set_parameters(AlignmentParameters *parameters) { _parameters = parameters; }

extern "C" void set_logger(AlignmentLogger *logger) { _logger = logger; }

extern "C" void delete_alignment_kernel(AlignmentKernel *instance) {
    if (instance != 0) delete instance;
}
