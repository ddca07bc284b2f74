/* versalign_fasta.h -- FASTA ingest straight into the packed layout the batch-friendly entry points
 * take (versalign_cuda.h: va_cuda_score_packed / va_cuda_align_packed).  Host code only; exported by
 * libCUDAKernel.so next to the kernels' C ABI.
 *
 * What it replaces in the reference: FastaProvider::parse_fasta (src/util/versalignUtil.h:53-93), which
 * strdup()s every record, followed by pad() (src/util/versalignUtil.cpp:17-33), which copies every
 * record once more into a heap block padded to the batch maximum -- 2n heap blocks that the kernels'
 * callers then hand over pointer by pointer.  Here a file becomes ONE block of bases plus n+1 offsets.
 *
 * Record rules are the reference parser's, quirks included (tests pin them against the reference's
 * own code, oracle/ref_util_shim.cpp):
 *   - only lines terminated by '\n' count (its loop is `while (getline(in, line).good())`: an
 *     unterminated last line is dropped);
 *   - a line that is empty or starts with '>' closes the current record; '>' + text opens a new one,
 *     a bare '>' or an empty line opens none, and sequence lines outside a record are ignored;
 *   - a sequence line containing a space discards the record it belongs to;
 *   - lines are concatenated as they are ('\r' stays, case stays); a record ends at its first NUL
 *     byte (the reference strdup()s a c_str());
 *   - an unreadable file is an error here (the reference prints a message and returns no records).
 */
#ifndef VERSALIGN_FASTA_H
#define VERSALIGN_FASTA_H

#include <stddef.h>
#include <stdint.h>

#include "versalign_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* *bases = alloc(total bytes, user): all records back to back, no terminators;
 * *offsets = (int64_t*)alloc((n+1)*8, user): record i = (*bases)[(*offsets)[i] .. (*offsets)[i+1]);
 * *n_records, *max_length (the value pad() would return) are plain outputs.
 * Returns VA_OK, VA_ERR_ARG (null argument / file cannot be read) or VA_ERR_MEMORY. */
int va_fasta_load(const char *path, va_cuda_alloc_fn alloc, void *user,
                  char **bases, int64_t **offsets, int64_t *n_records, int64_t *max_length);

#ifdef __cplusplus
}
#endif
#endif /* VERSALIGN_FASTA_H */
