"""FASTA ingest into the packed layout (include/versalign_fasta.h) against the reference's own
FastaProvider::parse_fasta, compiled from /root/reference into oracle/_ref/libref_util.so
(oracle/ref_util_shim.cpp + oracle/Makefile).  Host code: runs without a GPU."""
import os

import numpy as np
import pytest

from oracle import binding as ora
from versalignlib_b200 import capi, synth


def _records(bases, offsets):
    return [bases[offsets[i]:offsets[i + 1]].tobytes() for i in range(len(offsets) - 1)]


def _python_rules(text: bytes) -> list[bytes]:
    """The reference parser's rules restated (versalignUtil.h:53-93) -- the fallback checker when the
    shim library is not there, and a cross-check of the shim when it is."""
    out, name, content = [], b"", b""
    lines = text.split(b"\n")[:-1]  # only '\n'-terminated lines are seen
    for line in lines:
        if not line or line[:1] == b">":
            if name:
                out.append(content.split(b"\0")[0])
                name = b""
            if line:
                name = line[1:]
            content = b""
        elif name:
            if b" " in line:
                name, content = b"", b""
            else:
                content += line
    if name:
        out.append(content.split(b"\0")[0])
    return out


CASES = {
    "plain": b">r1\nACGT\n>r2\nGGCC\nTTAA\n",
    "multiline_and_blank": b">a desc with spaces\nACGT\nACGT\n\n>b\nNNNN\n\nACGT\n>c\n",
    "unterminated_last_line": b">a\nACGT\n>b\nGGGG",
    "unterminated_header": b">a\nACGT\n>b",
    "space_in_sequence": b">a\nAC GT\nTTTT\n>b\nCCCC\n",
    "bare_header": b">\nACGT\n>x\nTT\n",
    "crlf": b">a\r\nACGT\r\nGG\r\n>b\r\nTT\r\n",
    "leading_garbage": b"ACGT\nGGGG\n>a\nCC\n",
    "lower_and_n": b">a\nacgtNNacgt\n>b\nRYKM\n",
    "empty_record": b">a\n>b\nACGT\n",
    "empty_file": b"",
    "only_newlines": b"\n\n\n",
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_rules_match_the_reference_parser(tmp_path, name):
    path = str(tmp_path / (name + ".fa"))
    with open(path, "wb") as f:
        f.write(CASES[name])
    bases, offsets, max_len = capi.fasta_load(path)
    got = _records(bases, offsets)
    assert got == _python_rules(CASES[name])
    ref = ora.ref_parse_fasta(path)
    if ref is not None:  # pinned by execution when the reference was compiled here
        assert got == ref
    assert max_len == max([len(r) for r in got], default=0)


def test_large_file_round_trip(tmp_path):
    """20 k records of mixed length, 60 columns per line: the packed layout holds exactly what the
    reference parser returns, and it is what pack_batch makes of the padded batch."""
    reads, refs, rl, fl = synth.mixed_batch(20_000, 30, 250, seed=11)
    path = str(tmp_path / "reads.fa")
    with open(path, "wb") as f:
        for i in range(reads.shape[0]):
            seq = reads[i, :rl[i]].tobytes()
            f.write(b">read%d len=%d\n" % (i, rl[i]))
            for o in range(0, len(seq), 60):
                f.write(seq[o:o + 60] + b"\n")
    bases, offsets, max_len = capi.fasta_load(path)
    want_bases, want_off = synth.pack_batch(reads, rl)
    assert np.array_equal(offsets, want_off) and np.array_equal(bases, want_bases)
    assert max_len == int(rl.max())
    ref = ora.ref_parse_fasta(path)
    if ref is not None:
        assert _records(bases, offsets) == ref


def test_unreadable_file_is_an_error(tmp_path):
    with pytest.raises(capi.CudaError):
        capi.fasta_load(str(tmp_path / "missing.fa"))
