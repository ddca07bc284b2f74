// plugin_host.cpp -- driver-side loader for versalignLib kernel plug-ins, with a C ABI
// so Python (ctypes) and the C++ bench tool can drive ANY AlignmentKernel library --
// the reference's libDefaultKernel/libSSEKernel/libAVXKernel or our libCUDAKernel --
// through exactly the calls the reference driver makes:
//
//   reference                                           here
//   versalignUtil.cpp:45-76  DLL_init: dlopen(RTLD_LAZY) -> set_parameters -> set_logger   vah_load
//   main.cpp:227-238         get_kernel: dlsym spawn_alignment_kernel                      vah_load
//   main.cpp:217-225         clear_kernel: dlsym delete_alignment_kernel                   vah_close
//   CustomParameters.h:9-58  key -> int provider, throws on unknown key                    HostParameters
//   CustomLogger.h:19-59     stderr logger                                                 HostLogger
//   versalignUtil.cpp:17-33  pad(): one heap block of exactly max_length bytes per seq     vah_stage
//   main.cpp:131,143         kernel->score_alignments / compute_alignments                 vah_score / vah_align
//
// No DP arithmetic lives here.
#include "versalign_plugin_abi.h"

#include <dlfcn.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <map>
#include <string>
#include <vector>

namespace {

class HostParameters : public AlignmentParameters {
public:
    std::map<std::string, int> kv;
    int param_int(char const *const key) override {
        auto it = kv.find(key);
        if (it == kv.end()) {
            // same convention as the reference driver: unknown key throws a C string
            static thread_local std::string msg;
            msg = std::string("Unknown int parameter: ") + key;
            throw msg.c_str();
        }
        return it->second;
    }
    bool has_key(char const *const key) override { return kv.count(key) != 0; }
};

class HostLogger : public AlignmentLogger {
public:
    int verbosity = 1;  // 0: silent, 1: warnings and worse, 2: everything
    long counts[4] = {0, 0, 0, 0};
    std::string last_severe;
    void log(int const level, char const *const main, char const *const msg, size_t const &arg_num = 0,
             ...) override {
        (void)arg_num;
        int bucket = level == 0 ? 0 : level == 1 ? 1 : level == 3 ? 3 : 2;
        counts[bucket]++;
        if (bucket >= 2) last_severe = std::string("[") + main + "] " + msg;
        bool show = verbosity >= 2 || (verbosity == 1 && bucket >= 1);
        if (show) {
            const char *sev = bucket == 0 ? "INFO" : bucket == 1 ? "WARNING" : bucket == 3 ? "DRASTIC" : "ERROR";
            fprintf(stderr, "%s\t[%s]\t%s\n", sev, main, msg);
        }
    }
};

struct Host {
    HostParameters params;
    HostLogger logger;
    void *dll = nullptr;
    AlignmentKernel *kernel = nullptr;
    fp_delete_alignment_kernel destroy = nullptr;
    void (*set_parameters)(AlignmentParameters *) = nullptr;
    std::string error;

    // staged batch: what the reference driver holds after parse_fasta + pad()
    int n = 0, read_length = 0, ref_length = 0;
    std::vector<char *> reads, refs;
    bool scattered = false;
    std::vector<char> flat_reads, flat_refs;

    // result of the last compute_alignments call
    Alignment *alignments = nullptr;
    int n_alignments = 0;
    double last_call_seconds = 0.0;

    void drop_stage() {
        if (scattered) {
            for (char *p : reads) delete[] p;
            for (char *p : refs) delete[] p;
        }
        reads.clear();
        refs.clear();
        flat_reads.clear();
        flat_refs.clear();
        n = 0;
    }
    void drop_alignments() {
        delete[] alignments;  // runs ~Alignment -> delete[] read / ref
        alignments = nullptr;
        n_alignments = 0;
    }
};

int fail(Host *h, const std::string &msg) {
    h->error = msg;
    return -1;
}

}  // namespace

extern "C" {

void *vah_create(void) { return new Host(); }

void vah_set_param(void *hv, const char *key, int value) { static_cast<Host *>(hv)->params.kv[key] = value; }

void vah_unset_param(void *hv, const char *key) { static_cast<Host *>(hv)->params.kv.erase(key); }

void vah_set_verbosity(void *hv, int v) { static_cast<Host *>(hv)->logger.verbosity = v; }

const char *vah_error(void *hv) { return static_cast<Host *>(hv)->error.c_str(); }

long vah_log_count(void *hv, int bucket) {
    return (bucket >= 0 && bucket < 4) ? static_cast<Host *>(hv)->logger.counts[bucket] : -1;
}

// DLL_init + get_kernel.  Returns 0 on success.
int vah_load(void *hv, const char *path) {
    Host *h = static_cast<Host *>(hv);
    if (h->kernel) return fail(h, "a kernel is already loaded");
    h->dll = dlopen(path, RTLD_LAZY);
    if (!h->dll) return fail(h, std::string("Failed loading DLL: ") + path + ": " + dlerror());
    h->set_parameters = reinterpret_cast<void (*)(AlignmentParameters *)>(dlsym(h->dll, "set_parameters"));
    if (h->set_parameters) h->set_parameters(&h->params);
    auto set_logger = reinterpret_cast<void (*)(AlignmentLogger *)>(dlsym(h->dll, "set_logger"));
    if (set_logger) set_logger(&h->logger);
    auto spawn = reinterpret_cast<fp_load_alignment_kernel>(dlsym(h->dll, "spawn_alignment_kernel"));
    h->destroy = reinterpret_cast<fp_delete_alignment_kernel>(dlsym(h->dll, "delete_alignment_kernel"));
    if (!spawn || !h->destroy) return fail(h, "COULD NOT FIND FUNCTION spawn_alignment_kernel/delete_alignment_kernel");
    try {
        h->kernel = spawn();
    } catch (const char *msg) {  // the kernels throw C strings from their constructors
        return fail(h, std::string("spawn_alignment_kernel threw: ") + msg);
    } catch (...) {
        return fail(h, "spawn_alignment_kernel threw");
    }
    if (!h->kernel) return fail(h, "spawn_alignment_kernel returned null");
    return 0;
}

// main.cpp:261-265: change parameters, then spawn a fresh instance from the same library.
int vah_respawn(void *hv) {
    Host *h = static_cast<Host *>(hv);
    if (!h->dll) return fail(h, "no library loaded");
    if (h->kernel) {
        h->destroy(h->kernel);
        h->kernel = nullptr;
    }
    if (h->set_parameters) h->set_parameters(&h->params);
    auto spawn = reinterpret_cast<fp_load_alignment_kernel>(dlsym(h->dll, "spawn_alignment_kernel"));
    try {
        h->kernel = spawn();
    } catch (const char *msg) {
        return fail(h, std::string("spawn_alignment_kernel threw: ") + msg);
    } catch (...) {
        return fail(h, "spawn_alignment_kernel threw");
    }
    return h->kernel ? 0 : fail(h, "spawn_alignment_kernel returned null");
}

// Stage a batch the way the reference driver holds it: n sequences, each exactly
// read_length / ref_length bytes ('\0' padded by the caller), not NUL terminated.
// scattered != 0: one heap block per sequence like pad(); else pointers into one copy.
int vah_stage(void *hv, int n, const char *reads_flat, int read_length, const char *refs_flat, int ref_length,
              int scattered) {
    Host *h = static_cast<Host *>(hv);
    h->drop_stage();
    h->n = n;
    h->read_length = read_length;
    h->ref_length = ref_length;
    h->scattered = scattered != 0;
    h->reads.resize(n);
    h->refs.resize(n);
    if (h->scattered) {
        for (int i = 0; i < n; ++i) {  // interleaved, so read i and ref i are not adjacent to read i+1
            h->reads[i] = new char[read_length > 0 ? read_length : 1];
            memcpy(h->reads[i], reads_flat + (size_t)i * read_length, read_length);
            h->refs[i] = new char[ref_length > 0 ? ref_length : 1];
            memcpy(h->refs[i], refs_flat + (size_t)i * ref_length, ref_length);
        }
    } else {
        h->flat_reads.assign(reads_flat, reads_flat + (size_t)n * read_length);
        h->flat_refs.assign(refs_flat, refs_flat + (size_t)n * ref_length);
        for (int i = 0; i < n; ++i) {
            h->reads[i] = h->flat_reads.data() + (size_t)i * read_length;
            h->refs[i] = h->flat_refs.data() + (size_t)i * ref_length;
        }
    }
    return 0;
}

double vah_last_call_seconds(void *hv) { return static_cast<Host *>(hv)->last_call_seconds; }

// kernel->score_alignments(opt, n, reads, refs, scores) on the staged batch.
int vah_score(void *hv, int opt, int16_t *scores) {
    Host *h = static_cast<Host *>(hv);
    if (!h->kernel) return fail(h, "no kernel");
    try {
        auto t0 = std::chrono::steady_clock::now();
        h->kernel->score_alignments(opt, h->n, h->reads.data(), h->refs.data(), scores);
        h->last_call_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } catch (const char *msg) {
        return fail(h, std::string("score_alignments threw: ") + msg);
    } catch (...) {
        return fail(h, "score_alignments threw: " + h->logger.last_severe);
    }
    return 0;
}

// kernel->compute_alignments(opt, n, reads, refs, alignments) on the staged batch into a
// fresh value-initialised Alignment[n] (main.cpp:123).  Results stay inside the host
// object until vah_fetch_alignments / the next call.
int vah_align(void *hv, int opt) {
    Host *h = static_cast<Host *>(hv);
    if (!h->kernel) return fail(h, "no kernel");
    h->drop_alignments();
    h->alignments = new Alignment[h->n]();
    h->n_alignments = h->n;
    try {
        auto t0 = std::chrono::steady_clock::now();
        h->kernel->compute_alignments(opt, h->n, h->reads.data(), h->refs.data(), h->alignments);
        h->last_call_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } catch (const char *msg) {
        return fail(h, std::string("compute_alignments threw: ") + msg);
    } catch (...) {
        return fail(h, "compute_alignments threw: " + h->logger.last_severe);
    }
    return 0;
}

// Flatten the last compute_alignments result: per pair aln_length bytes of each string
// (bytes before `start` are zeroed here -- the kernels leave them undefined), and the four
// shorts.  A pair whose kernel left read/ref null gets start = -1.
int vah_fetch_alignments(void *hv, char *aln_read, char *aln_ref, int16_t *fields /* n*4 */) {
    Host *h = static_cast<Host *>(hv);
    if (!h->alignments) return fail(h, "no alignments");
    const int L = h->read_length + h->ref_length;
    for (int i = 0; i < h->n_alignments; ++i) {
        const Alignment &a = h->alignments[i];
        char *orow = aln_read + (size_t)i * L, *frow = aln_ref + (size_t)i * L;
        memset(orow, 0, L);
        memset(frow, 0, L);
        if (!a.read || !a.ref) {
            fields[4 * i] = fields[4 * i + 1] = fields[4 * i + 2] = fields[4 * i + 3] = -1;
            continue;
        }
        int s = a.readStart;
        if (s < 0) s = 0;
        if (s < L - 1) {
            memcpy(orow + s, a.read + s, (size_t)(L - 1 - s));
            memcpy(frow + s, a.ref + s, (size_t)(L - 1 - s));
        }
        fields[4 * i] = a.readStart;
        fields[4 * i + 1] = a.readEnd;
        fields[4 * i + 2] = a.refStart;
        fields[4 * i + 3] = a.refEnd;
    }
    return 0;
}

// 1 if byte aln_length-1 of both strings is NUL for every pair (the consumers print
// read+readStart as a C string, main.cpp:147-153).
int vah_alignments_terminated(void *hv) {
    Host *h = static_cast<Host *>(hv);
    if (!h->alignments) return -1;
    const int L = h->read_length + h->ref_length;
    for (int i = 0; i < h->n_alignments; ++i) {
        const Alignment &a = h->alignments[i];
        if (!a.read || !a.ref || a.read[L - 1] != 0 || a.ref[L - 1] != 0) return 0;
    }
    return 1;
}

void vah_drop_alignments(void *hv) { static_cast<Host *>(hv)->drop_alignments(); }

void vah_close(void *hv) {
    Host *h = static_cast<Host *>(hv);
    h->drop_alignments();
    h->drop_stage();
    if (h->kernel && h->destroy) h->destroy(h->kernel);
    // like the reference driver, never dlclose(): plug-ins keep static state
    delete h;
}

}  // extern "C"
