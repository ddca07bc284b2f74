"""Driver / benchmark harness for AlignmentKernel plug-ins.

What it replaces in the reference: main() and time_kernel() of src/impl/main.cpp (:74-215, :240-295),
which hard-code the kernel name, the test-set paths, SW alignment and 100 repetitions, and leak every
result.  Same plug-in boundary (dlopen -> set_parameters -> set_logger -> spawn_alignment_kernel -> the
two virtual calls, through csrc/plugin_host.cpp), but every choice is an argument, any kernel library can
be put next to any other for a parity verdict, and the outcome is one JSON object.

  python -m versalignlib_b200.driver --mode nw_align --synthetic 100000,150,150
  python -m versalignlib_b200.driver --mode sw_score --reads reads.fa --refs refs.fa \\
         --compare /path/to/libDefaultKernel.so --reps 5
  python -m versalignlib_b200.driver --kernel /path/to/libSSEKernel.so --threads 1 --mode sw_score ...

--kernel / --compare take the path of any library exporting the reference's four plug-in symbols
(default --kernel: this package's libCUDAKernel.so).  Sequences come from two FASTA files (record i of
one against record i of the other; parsed by va_fasta_load and padded the way the reference's pad() does)
or from the seeded generator.  GCUPS counts read_length x ref_length cells per pair, as the reference's
kernels compute them (padded).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

import numpy as np

from . import capi, synth
from .host import NW, SW, PluginHost

MODES = {"sw_score": (SW, False), "nw_score": (NW, False), "sw_align": (SW, True), "nw_align": (NW, True)}


def padded_from_fasta(path: str, limit: int | None):
    bases, off, _ = capi.fasta_load(path)
    n = len(off) - 1 if limit is None else min(limit, len(off) - 1)
    lens = np.diff(off[: n + 1])
    width = int(lens.max()) if n else 0
    out = np.zeros((n, width), dtype=np.uint8)
    col = np.arange(width)[None, :]
    out[col < lens[:, None]] = bases[: int(off[n])]
    return out


def run(library: str, opt: int, align: bool, reads, refs, scoring, threads: int, reps: int, extra: dict):
    """Returns (result arrays, per-call seconds)."""
    times, result = [], None
    with PluginHost(library, reads.shape[1], refs.shape[1], scoring, num_threads=threads, extra=extra, verbosity=0) as h:
        h.stage(reads, refs, scattered=True)
        for rep in range(reps + 1):  # one warm-up
            if align:
                h.align_staged(opt, fetch=False)
                if rep == reps:
                    result = h.fetch_alignments()
                h.drop_alignments()
            else:
                result = (h.score_staged(opt),)
            if rep:
                times.append(h.last_call_seconds)
    return result, times


def verdict(a, b, align: bool, L: int) -> dict:
    if not align:
        bad = np.nonzero(a[0] != b[0])[0]
        return {"compared": "scores", "pairs": int(a[0].shape[0]), "mismatches": int(bad.size), "first": bad[:5].tolist()}
    (ar, af, fa), (br, bf, fb) = a, b
    start_a, start_b = fa[:, 0].astype(np.int64), fb[:, 0].astype(np.int64)
    col = np.arange(L)[None, :]
    used = (col >= np.clip(start_a, 0, L)[:, None]) & (col < L - 1)  # bytes before start are undefined in the reference
    bad = (start_a != start_b) | ((ar != br) & used).any(axis=1) | ((af != bf) & used).any(axis=1)
    bad = np.nonzero(bad)[0]
    return {"compared": "start offsets + used bytes of both gapped strings", "pairs": int(ar.shape[0]),
            "mismatches": int(bad.size), "first": bad[:5].tolist()}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m versalignlib_b200.driver", description=__doc__.split("\n\n")[0])
    ap.add_argument("--kernel", default=None, help="plug-in library to run (default: libCUDAKernel.so of this package)")
    ap.add_argument("--compare", default=None, help="second plug-in library: run it on the same input and report parity")
    ap.add_argument("--mode", choices=sorted(MODES), default="sw_score")
    ap.add_argument("--reads", help="FASTA file of reads")
    ap.add_argument("--refs", help="FASTA file of references (record i pairs with read i)")
    ap.add_argument("--synthetic", help="n,read_length,ref_length : seeded synthetic pairs instead of files")
    ap.add_argument("--limit", type=int, default=None, help="use only the first N pairs")
    ap.add_argument("--scoring", default="2,-1,-3,-3", help="match,mismatch,gap_read,gap_ref")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1, help="num_threads parameter of the kernels")
    ap.add_argument("--compare-threads", type=int, default=None, help="num_threads for --compare (the reference SSE kernel needs 1)")
    ap.add_argument("--policy", type=int, default=0, help="cuda_traceback_policy: 0 Default/OpenCL rule, 1 SSE/AVX rule")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args(argv)

    opt, align = MODES[a.mode]
    scoring = tuple(int(x) for x in a.scoring.split(","))
    if a.synthetic:
        n, rl, fl = (int(x) for x in a.synthetic.split(","))
        reads, refs = synth.uniform_batch(n, rl, fl, p_sub=0.08, q_indel=0.02 if align else 0.0, seed=synth.BASE_SEED + 2)
        source = f"synthetic {n} x ({rl} vs {fl})"
    elif a.reads and a.refs:
        reads, refs = padded_from_fasta(a.reads, a.limit), padded_from_fasta(a.refs, a.limit)
        n = min(reads.shape[0], refs.shape[0])
        reads, refs = np.ascontiguousarray(reads[:n]), np.ascontiguousarray(refs[:n])
        source = f"{a.reads} x {a.refs}"
    else:
        ap.error("give --reads and --refs, or --synthetic n,read_length,ref_length")
    if a.limit is not None:
        reads, refs = np.ascontiguousarray(reads[: a.limit]), np.ascontiguousarray(refs[: a.limit])
    n, rl, fl = reads.shape[0], reads.shape[1], refs.shape[1]
    cells = float(n) * rl * fl

    kernel = a.kernel or capi.library_path()
    ours = os.path.samefile(kernel, capi.library_path()) if os.path.exists(kernel) else False
    extra = {"cuda_traceback_policy": a.policy} if ours else {}
    result, times = run(kernel, opt, align, reads, refs, scoring, a.threads, a.reps, extra)
    sec = statistics.median(times)
    out = {"kernel": kernel, "mode": a.mode, "input": source, "pairs": n, "read_length": rl, "ref_length": fl,
           "scoring": list(scoring), "threads": a.threads, "reps": a.reps, "seconds_median": sec,
           "seconds_all": times, "gcups": cells / sec / 1e9}
    if a.compare:
        cthreads = a.compare_threads or a.threads
        other, ctimes = run(a.compare, opt, align, reads, refs, scoring, cthreads, 1, {})
        csec = statistics.median(ctimes)
        out["compare"] = {"kernel": a.compare, "threads": cthreads, "seconds": csec, "gcups": cells / csec / 1e9,
                          "speedup": csec / sec, "parity": verdict(result, other, align, rl + fl)}
    print(json.dumps(out))
    return 0 if not a.compare or out["compare"]["parity"]["mismatches"] == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
