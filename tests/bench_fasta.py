"""Host-side throughput of the FASTA ingest (va_fasta_load -> packed layout) next to the reference's
parse_fasta (+ the pad() copy it needs before a kernel call), on a synthetic file.
usage: python tests/bench_fasta.py [records] [length]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import binding as ora  # noqa: E402
from versalignlib_b200 import capi, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    length = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    seqs = synth.random_seqs(np.random.Generator(np.random.PCG64(1)), n, length)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "reads.fa")
        with open(path, "wb") as f:
            lines = [b">r%d\n%s\n" % (i, seqs[i].tobytes()) for i in range(n)]
            f.write(b"".join(lines))
        size = os.path.getsize(path)
        t0 = time.perf_counter()
        bases, off, mx = capi.fasta_load(path)
        t1 = time.perf_counter()
        out = {"records": n, "length": length, "file_MB": round(size / 1e6, 1),
               "va_fasta_load_s": round(t1 - t0, 3), "va_fasta_load_MBps": round(size / 1e6 / (t1 - t0), 1)}
        t0 = time.perf_counter()
        ref = ora.ref_parse_fasta(path)
        t1 = time.perf_counter()
        if ref is not None:
            out["reference_parse_fasta_s"] = round(t1 - t0, 3)
            out["reference_parse_fasta_MBps"] = round(size / 1e6 / (t1 - t0), 1)
            out["note"] = "reference time includes copying its strings into Python; pad() would add one more heap block per record"
        print(out)


if __name__ == "__main__":
    main()
