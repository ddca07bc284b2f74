// Test infrastructure only: exposes the reference's own FASTA parser and pad() (compiled from
// /root/reference where they lie -- see oracle/Makefile, target _ref/libref_util.so) through a C ABI so
// the tests can pin va_fasta_load against them.  Nothing of the reference is copied here.
#include <cstdlib>
#include <cstring>

#include "versalignUtil.h"  // FastaProvider (src/util/versalignUtil.h:47-93), pad() (versalignUtil.cpp:17-33)

extern "C" {

// returns the number of records; *seqs = malloc'ed array of the parser's strdup'ed strings
int ref_parse_fasta(const char *path, char ***seqs) {
    FastaProvider provider;
    std::vector<const char *> v = provider.parse_fasta(std::string(path));
    char **out = (char **)malloc(sizeof(char *) * (v.size() ? v.size() : 1));
    for (size_t i = 0; i < v.size(); ++i) out[i] = const_cast<char *>(v[i]);
    *seqs = out;
    return (int)v.size();
}

// pads the strings in place like the driver does (main.cpp:106-110) and returns the padded length
size_t ref_pad(const char **strings, int n) { return pad(strings, n, '\0'); }

void ref_free_strings(char **seqs, int n) {
    for (int i = 0; i < n; ++i) free(seqs[i]);
    free(seqs);
}
}
