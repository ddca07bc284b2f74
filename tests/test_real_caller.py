"""The reference's REAL caller: src/impl/main.cpp built into oracle/_ref/Versalign (oracle/Makefile documents the
four sed edits: register "CUDA" in kernel_map, kernel name from $VERSALIGN_KERNEL, drop entries whose library
is absent, num_loops as a -D).  The executable finds its libraries under ../bin/Versalign-0.1.0/lib and its input
under ../testset relative to the working directory (main.cpp:29-33,88-93) and writes four text files
(main.cpp:131-184); the tests lay that tree out in a scratch directory.

  CPU : the `default` kernel through the real caller reproduces the oracle (pins the oracle a second way)
  GPU : libCUDAKernel.so through the real caller -- dlopen, set_parameters, spawn, the two virtual calls, the
        timing loop that re-spawns kernels under changing num_threads -- writes the same four files as `default`
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import binding as ora
from versalignlib_b200 import build as vbuild
from versalignlib_b200 import synth

VERSALIGN = os.path.join(ora.REF_DIR, "Versalign")
SCORING = (2, -1, -3, -3)  # CustomParameters.h defaults

pytestmark = pytest.mark.skipif(not os.path.exists(VERSALIGN) or ora.ref_lib("Default") is None,
                                reason="reference caller / kernels not built (oracle/_ref)")


def _layout(tmp_path, libs, n=1008):
    """scratch tree: run/ (cwd), bin/Versalign-0.1.0/lib/, testset/{reads,refs}.fa; returns the padded batch"""
    lib_dir = tmp_path / "bin" / "Versalign-0.1.0" / "lib"
    lib_dir.mkdir(parents=True)
    for name, path in libs.items():
        shutil.copy(path, lib_dir / name)
    (tmp_path / "testset").mkdir()
    (tmp_path / "run").mkdir()
    reads, refs, rl, fl = synth.mixed_batch(n, 40, 120, p_sub=0.1, seed=77)
    for name, arr, lens in (("reads.fa", reads, rl), ("refs.fa", refs, fl)):
        with open(tmp_path / "testset" / name, "wb") as f:
            for i in range(n):
                f.write(b">s%d\n%s\n" % (i, arr[i, :lens[i]].tobytes()))
    return reads, refs


def _run(tmp_path, kernel):
    env = dict(os.environ, VERSALIGN_KERNEL=kernel, OMP_NUM_THREADS="4")
    r = subprocess.run([VERSALIGN], cwd=tmp_path / "run", env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = {}
    for name in ("scores_smith_waterman", "alignments_smith_waterman", "scores_needleman_wunsch", "alignments_needleman_wunsch"):
        with open(tmp_path / "run" / f"{name}.txt", "rb") as f:
            out[name] = f.read()
    return out, r.stdout


def _scores(blob):
    return np.array([int(l.rsplit(b"\t", 1)[1]) for l in blob.splitlines() if b"\t" in l], dtype=np.int64)


def _records(blob):
    """alignment files: read line, ref line, empty line per pair"""
    lines = blob.split(b"\n")
    return [(lines[i], lines[i + 1]) for i in range(0, len(lines) - 2, 3)]


def _strings(a, b, start):
    """what main.cpp:151-152 prints: the block from `start` up to the first NUL"""
    return [(bytes(a[i, start[i]:]).split(b"\0", 1)[0], bytes(b[i, start[i]:]).split(b"\0", 1)[0]) for i in range(a.shape[0])]


def test_default_kernel_through_real_caller_matches_oracle(tmp_path):
    reads, refs = _layout(tmp_path, {"libDefaultKernel.so": ora.ref_lib("Default")})
    out, stdout = _run(tmp_path, "default")
    n = reads.shape[0]
    # Default stores only the low byte of a score (DefaultKernel.cpp:137,199); the caller's array starts zeroed
    assert np.array_equal(_scores(out["scores_smith_waterman"]), ora.score(ora.SW, reads, refs, SCORING).astype(np.int64) & 0xFF)
    assert np.array_equal(_scores(out["scores_needleman_wunsch"]), ora.score(ora.NW, reads, refs, SCORING).astype(np.int64) & 0xFF)
    for mode, name in ((ora.NW, "alignments_needleman_wunsch"), (ora.SW, "alignments_smith_waterman")):
        a, b, start, _ = ora.align(mode, ora.POLICY_DEFAULT_OCL, reads, refs, SCORING)
        got, want = _records(out[name]), _strings(a, b, start)
        assert len(got) == n
        # Default's SW traceback never writes the terminating NUL (DefaultKernel.cpp:391-456): its lines may carry
        # heap bytes after the alignment, so the oracle's string must be a prefix there and equal for NW
        for g, w in zip(got, want):
            assert (g[0].startswith(w[0]) and g[1].startswith(w[1])) if mode == ora.SW else g == w
    assert "default" in stdout and "Threads" in stdout


@pytest.mark.gpu
def test_cuda_kernel_through_real_caller(tmp_path):
    reads, refs = _layout(tmp_path, {"libDefaultKernel.so": ora.ref_lib("Default"), "libCUDAKernel.so": vbuild.CUDA_PLUGIN})
    ref_out, _ = _run(tmp_path, "default")
    out, stdout = _run(tmp_path, "CUDA")
    n = reads.shape[0]
    # scores: the CUDA kernel writes the full short; every score of this deck is below 256, so the files agree
    for name in ("scores_smith_waterman", "scores_needleman_wunsch"):
        assert np.array_equal(_scores(out[name]), _scores(ref_out[name]))
    assert np.array_equal(_scores(out["scores_smith_waterman"]), ora.score(ora.SW, reads, refs, SCORING).astype(np.int64))
    assert out["alignments_needleman_wunsch"] == ref_out["alignments_needleman_wunsch"]
    got, want = _records(out["alignments_smith_waterman"]), _records(ref_out["alignments_smith_waterman"])
    assert len(got) == n == len(want)
    for g, w in zip(got, want):  # see above: Default's SW lines may carry trailing heap bytes
        assert w[0].startswith(g[0]) and w[1].startswith(g[1])
    a, b, start, _ = ora.align(ora.SW, ora.POLICY_DEFAULT_OCL, reads, refs, SCORING)
    assert got == _strings(a, b, start)
    # the timing section (main.cpp:186-203, time_kernel) re-spawned the CUDA kernel under 7 num_threads settings
    row = [l for l in stdout.splitlines() if l.startswith("CUDA")]
    assert row and len(row[0].split("\t")) == 8, stdout[-1000:]
