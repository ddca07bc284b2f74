// versalign_plugin_abi.h -- the C++ side of the versalignLib kernel plug-in boundary,
// restated in one header so that libCUDAKernel.so can be built without the
// reference tree.  A host that was compiled against the reference's own three
// interface headers can load a library compiled against this one and vice versa:
// what matters across dlopen() is the Itanium C++ ABI layout, i.e. the ORDER of the
// virtual members and the layout of the Alignment record, both pinned below.
//
//   reference header                      what it fixes                       here
//   include/AlignmentKernel.h:12-24       result record + array-delete dtor   struct Alignment
//   include/AlignmentKernel.h:34-44       vtable: dtor, score_*, compute_*    class AlignmentKernel
//   include/AlignmentKernel.h:46-47       factory function pointer types      fp_load_/fp_delete_
//   include/AlignmentParameters.h:11-22   vtable: param_int, has_key, dtor    class AlignmentParameters
//   include/AlignmentLogger.h:13-22       vtable: log, dtor                   class AlignmentLogger
//
// If the reference headers are on the include path first they win (same guards).
#ifndef VERSALIGN_PLUGIN_ABI_H
#define VERSALIGN_PLUGIN_ABI_H

#include <stddef.h>

// ---- AlignmentKernel.h ----------------------------------------------------------
#ifndef ALIGNMENTKERNEL_H
#define ALIGNMENTKERNEL_H

// Result of compute_alignments for one pair.  `read` and `ref` are two gapped
// strings of the same length, each in its own heap block obtained with new char[]
// (the record's destructor releases them with delete[], so the plug-in must use
// array new).  The four shorts are OFFSETS INTO THOSE BLOCKS, not sequence
// coordinates: the alignment occupies [readStart, readEnd) == [refStart, refEnd).
struct Alignment {
    char *read = 0;
    char *ref = 0;
    short readStart;
    short readEnd;
    short refStart;
    short refEnd;
    ~Alignment() {
        if (read != 0) delete[] read;
        if (ref != 0) delete[] ref;
    }
};

// opt & 0xF selects the recurrence: 0 = Smith-Waterman (local),
// 1 = the library's "Needleman-Wunsch" (see oracle/va_oracle.h for what it really is).
class AlignmentKernel {
public:
    virtual ~AlignmentKernel() {}
    virtual void score_alignments(int const &opt, int const &aln_number,
                                  char const *const *const reads,
                                  char const *const *const refs,
                                  short *const scores) = 0;
    virtual void compute_alignments(int const &opt, int const &aln_number,
                                    char const *const *const reads,
                                    char const *const *const refs,
                                    Alignment *const alignments) = 0;
};

typedef AlignmentKernel *(*fp_load_alignment_kernel)();
typedef void (*fp_delete_alignment_kernel)(AlignmentKernel *);
#endif  // ALIGNMENTKERNEL_H

// ---- AlignmentParameters.h ------------------------------------------------------
#ifndef INCLUDE_ALIGNMENTPARAMETERS_H
#define INCLUDE_ALIGNMENTPARAMETERS_H
// Key -> int provider owned by the host.  param_int() on an unknown key may throw
// (the reference driver's implementation does), so always ask has_key() first.
class AlignmentParameters {
public:
    virtual int param_int(char const *const key) = 0;
    virtual bool has_key(char const *const key) = 0;
    virtual ~AlignmentParameters() {}
};
typedef void (*fp_set_parameters)(AlignmentParameters const *);
// every plug-in defines its own copy of this pointer (C++ linkage, not extern "C")
extern AlignmentParameters *_parameters;
#define Parameters (*_parameters)
#endif  // INCLUDE_ALIGNMENTPARAMETERS_H

// ---- AlignmentLogger.h ----------------------------------------------------------
#ifndef ALIGNMENTLOGGER_H
#define ALIGNMENTLOGGER_H
// level: 0 info, 1 warning, 3 fatal, anything else error.  Not thread-safe in the
// reference driver: log from the calling thread only.
class AlignmentLogger {
public:
    virtual void log(int const level, char const *const main, char const *const msg,
                     size_t const &arg_num = 0, ...) = 0;
    virtual ~AlignmentLogger() {}
};
typedef void (*fp_set_logger)(AlignmentLogger const *);
extern AlignmentLogger *_logger;
#define Logger (*_logger)
#endif  // ALIGNMENTLOGGER_H

#if defined(__cplusplus) && __cplusplus >= 201103L && defined(__x86_64__)
static_assert(sizeof(Alignment) == 24, "Alignment must be 2 pointers + 4 shorts (AlignmentKernel.h:12-18)");
#endif

#endif  // VERSALIGN_PLUGIN_ABI_H
