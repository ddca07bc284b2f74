// va_fasta.cpp -- FASTA ingest into the packed layout (include/versalign_fasta.h).  Host code only.
// Same record rules as the reference's FastaProvider::parse_fasta (src/util/versalignUtil.h:53-93),
// restructured: the file is read in one piece and cut with memchr, a first pass sizes the result and a
// second copies the sequence lines straight into ONE caller-allocated block instead of one strdup() each,
// and nothing is padded (pad(), versalignUtil.cpp:17-33, is what
// the packed entry points make unnecessary).
#include "versalign_fasta.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

// whole file in one read
bool read_file(const char *path, std::vector<char> &out) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    bool ok = fseek(f, 0, SEEK_END) == 0;
    const long size = ok ? ftell(f) : -1;
    ok = ok && size >= 0 && fseek(f, 0, SEEK_SET) == 0;
    if (ok) {
        out.resize((size_t)size);
        ok = size == 0 || fread(out.data(), 1, (size_t)size, f) == (size_t)size;
    }
    fclose(f);
    return ok;
}

// One pass over the lines with the reference's record rules.  begin(): a record opens; line(p, len): a
// sequence line of the open record; drop(): the open record is discarded; end(): it is complete.
template <class Begin, class Line, class Drop, class End>
void scan(const char *p, const char *end, Begin begin, Line line, Drop drop, End finish) {
    bool open = false;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!nl) break;  // an unterminated last line is not seen by the reference's loop
        const size_t len = (size_t)(nl - p);
        if (len == 0 || p[0] == '>') {
            if (open) finish();
            open = len > 1;  // '>' followed by a name opens a record; a bare '>' or an empty line does not
            if (open) begin();
        } else if (open) {
            if (memchr(p, ' ', len)) {  // a space inside a sequence line discards the record
                drop();
                open = false;
            } else {
                line(p, len);
            }
        }
        p = nl + 1;
    }
    if (open) finish();
}

}  // namespace

extern "C" int va_fasta_load(const char *path, va_cuda_alloc_fn alloc, void *user, char **bases, int64_t **offsets,
                             int64_t *n_records, int64_t *max_length) {
    if (!path || !alloc || !bases || !offsets || !n_records) return VA_ERR_ARG;
    *bases = nullptr;
    *offsets = nullptr;
    *n_records = 0;
    if (max_length) *max_length = 0;
    std::vector<char> text;
    if (!read_file(path, text)) return VA_ERR_ARG;
    const char *t0 = text.data(), *t1 = t0 + text.size();

    // pass 1: how many records, how many bytes (an upper bound: records are cut at their first NUL in pass 2)
    int64_t n = 0, total = 0, cur = 0;
    scan(t0, t1, [&] { cur = 0; }, [&](const char *, size_t len) { cur += (int64_t)len; }, [&] { cur = 0; },
         [&] { ++n; total += cur; });
    char *b = alloc(total ? (size_t)total : 1, user);
    int64_t *o = (int64_t *)alloc((size_t)(n + 1) * sizeof(int64_t), user);
    if (!b || !o) return VA_ERR_MEMORY;

    // pass 2: copy the sequence lines straight into the block
    int64_t k = 0, pos = 0, rec = 0, longest = 0;
    o[0] = 0;
    scan(t0, t1, [&] { rec = pos; },
         [&](const char *p, size_t len) { memcpy(b + pos, p, len); pos += (int64_t)len; },
         [&] { pos = rec; },
         [&] {
             // the reference hands out strdup(content.c_str()): the record ends at its first NUL byte
             const void *nul = pos > rec ? memchr(b + rec, 0, (size_t)(pos - rec)) : nullptr;
             if (nul) pos = (int64_t)((const char *)nul - b);
             o[++k] = pos;
             longest = pos - rec > longest ? pos - rec : longest;
         });
    *bases = b;
    *offsets = o;
    *n_records = n;
    if (max_length) *max_length = longest;
    return VA_OK;
}
