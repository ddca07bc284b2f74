"""ctypes binding of oracle/libva_oracle.so -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libva_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

SW, NW = 0, 1
POLICY_DEFAULT_OCL, POLICY_SIMD = 0, 1


class Scoring(ctypes.Structure):
    _fields_ = [("match", ctypes.c_int32), ("mismatch", ctypes.c_int32),
                ("gap_read", ctypes.c_int32), ("gap_ref", ctypes.c_int32)]


def build(force: bool = False) -> None:
    """Compile the oracle (and the reference kernels when /root/reference is mounted)."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(os.path.join(HERE, f)) for f in ("va_oracle.c", "va_oracle_affine.c", "va_oracle.h")):
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/src/Kernels"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        L.va_oracle_score.restype = ctypes.c_int
        L.va_oracle_score.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_int, ctypes.POINTER(Scoring), ctypes.c_void_p, ctypes.c_int]
        L.va_oracle_align.restype = ctypes.c_int
        L.va_oracle_align.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(Scoring), ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.va_oracle_score_affine.restype = ctypes.c_int
        L.va_oracle_score_affine.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.POINTER(Scoring), ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.va_oracle_align_affine.restype = ctypes.c_int
        L.va_oracle_align_affine.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.POINTER(Scoring), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


def _chk(a: np.ndarray) -> np.ndarray:
    assert a.dtype == np.uint8 and a.ndim == 2 and a.flags.c_contiguous
    return a


def score(opt: int, reads: np.ndarray, refs: np.ndarray, scoring=(2, -1, -3, -3), threads: int = 0) -> np.ndarray:
    reads, refs = _chk(reads), _chk(refs)
    n = reads.shape[0]
    out = np.zeros(n, dtype=np.int16)
    sc = Scoring(*scoring)
    rc = lib().va_oracle_score(opt, n, reads.ctypes.data, reads.shape[1], refs.ctypes.data, refs.shape[1],
                               ctypes.byref(sc), out.ctypes.data, threads)
    if rc != 0:
        raise ValueError(f"va_oracle_score rc={rc}")
    return out


def align(opt: int, policy: int, reads: np.ndarray, refs: np.ndarray, scoring=(2, -1, -3, -3), threads: int = 0):
    """Returns (aln_read[n,L], aln_ref[n,L], start[n], end_cell[n,2])."""
    reads, refs = _chk(reads), _chk(refs)
    n = reads.shape[0]
    L = reads.shape[1] + refs.shape[1]
    a = np.zeros((n, L), dtype=np.uint8)
    b = np.zeros((n, L), dtype=np.uint8)
    start = np.zeros(n, dtype=np.int16)
    end = np.zeros((n, 2), dtype=np.int16)
    sc = Scoring(*scoring)
    rc = lib().va_oracle_align(opt, policy, n, reads.ctypes.data, reads.shape[1], refs.ctypes.data, refs.shape[1],
                               ctypes.byref(sc), a.ctypes.data, b.ctypes.data, start.ctypes.data, end.ctypes.data,
                               threads)
    if rc != 0:
        raise ValueError(f"va_oracle_align rc={rc}")
    return a, b, start, end


def score_affine(opt: int, reads: np.ndarray, refs: np.ndarray, scoring=(2, -1, -3, -3), gap_open: int = -5, threads: int = 0) -> np.ndarray:
    """Affine-gap variant (not in the reference; see va_oracle_affine.c)."""
    reads, refs = _chk(reads), _chk(refs)
    out = np.zeros(reads.shape[0], dtype=np.int16)
    sc = Scoring(*scoring)
    rc = lib().va_oracle_score_affine(opt, reads.shape[0], reads.ctypes.data, reads.shape[1], refs.ctypes.data, refs.shape[1],
                                      ctypes.byref(sc), gap_open, out.ctypes.data, threads)
    if rc != 0:
        raise ValueError(f"va_oracle_score_affine rc={rc}")
    return out


def align_affine(opt: int, reads: np.ndarray, refs: np.ndarray, scoring=(2, -1, -3, -3), gap_open: int = -5, threads: int = 0):
    """Returns (aln_read[n,L], aln_ref[n,L], start[n], end_cell[n,2], scores[n])."""
    reads, refs = _chk(reads), _chk(refs)
    n, L = reads.shape[0], reads.shape[1] + refs.shape[1]
    a = np.zeros((n, L), dtype=np.uint8)
    b = np.zeros((n, L), dtype=np.uint8)
    start = np.zeros(n, dtype=np.int16)
    end = np.zeros((n, 2), dtype=np.int16)
    scores = np.zeros(n, dtype=np.int16)
    sc = Scoring(*scoring)
    rc = lib().va_oracle_align_affine(opt, n, reads.ctypes.data, reads.shape[1], refs.ctypes.data, refs.shape[1], ctypes.byref(sc),
                                      gap_open, a.ctypes.data, b.ctypes.data, start.ctypes.data, end.ctypes.data, scores.ctypes.data, threads)
    if rc != 0:
        raise ValueError(f"va_oracle_align_affine rc={rc}")
    return a, b, start, end, scores


def ref_lib(name: str) -> str | None:
    """Path of a reference kernel built from /root/reference (Default, SSE, AVX) or None."""
    p = os.path.join(REF_DIR, f"lib{name}Kernel.so")
    return p if os.path.exists(p) else None


def ref_parse_fasta(path: str) -> list[bytes] | None:
    """The reference's own FastaProvider::parse_fasta (oracle/_ref/libref_util.so, built from
    /root/reference by oracle/Makefile) -- None when that library is not there."""
    p = os.path.join(REF_DIR, "libref_util.so")
    if not os.path.exists(p):
        return None
    L = ctypes.CDLL(p)
    L.ref_parse_fasta.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_char_p))]
    L.ref_parse_fasta.restype = ctypes.c_int
    L.ref_free_strings.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int]
    seqs = ctypes.POINTER(ctypes.c_char_p)()
    n = L.ref_parse_fasta(path.encode(), ctypes.byref(seqs))
    out = [seqs[i] for i in range(n)]  # c_char_p -> bytes (copied)
    L.ref_free_strings(seqs, n)
    return out
