"""ctypes binding of the flat C ABI declared in include/versalign_cuda.h.

The library is libCUDAKernel.so itself (the plug-in exports both boundaries).  There is no
fallback: if the shared object or a CUDA device is missing, calls raise.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

SW, NW = 0, 1
SW_AFFINE, NW_AFFINE = 2, 3  # affine-gap variants (not in the reference): see affine_opt()
POLICY_DEFAULT_OCL, POLICY_SIMD = 0, 1


def affine_opt(opt: int, gap_open: int) -> int:
    """VA_OPT_*_AFFINE | VA_OPT_GAP_OPEN(gap_open) of include/versalign_cuda.h: the affine-gap variant of mode `opt`
    (SW or NW) with a gap-open score <= 0; every entry point takes the result as its `opt`."""
    if gap_open > 0 or gap_open < -0xFFFF:
        raise ValueError("gap_open must be in [-65535, 0]")
    return (SW_AFFINE if (opt & 0xF) == SW else NW_AFFINE) | (((-gap_open) & 0xFFFF) << 8)

# every symbol include/versalign_cuda.h declares (tests check the .so exports all of them)
C_ABI_SYMBOLS = [
    "va_cuda_abi_version", "va_cuda_last_error", "va_cuda_device_count", "va_cuda_create", "va_cuda_destroy",
    "va_cuda_set_host_threads", "va_cuda_get_timings", "va_cuda_score_ptrs", "va_cuda_align_ptrs", "va_cuda_align_alloc", "va_cuda_align_records", "va_cuda_score_packed", "va_cuda_align_packed", "va_fasta_load",
    "va_cuda_score_flat", "va_cuda_align_flat", "va_cuda_score_device", "va_cuda_align_device",
    "va_cuda_max_resident_pairs", "va_cuda_int_peak", "va_cuda_set_profiling", "va_cuda_get_kernel_ms",
    "va_cuda_plugin_timings",
]
PLUGIN_SYMBOLS = ["spawn_alignment_kernel", "delete_alignment_kernel", "set_parameters", "set_logger"]


class Scoring(ctypes.Structure):
    _fields_ = [("match", ctypes.c_int32), ("mismatch", ctypes.c_int32),
                ("gap_read", ctypes.c_int32), ("gap_ref", ctypes.c_int32)]


class Timings(ctypes.Structure):
    _fields_ = [("total_s", ctypes.c_double), ("gather_s", ctypes.c_double), ("scatter_s", ctypes.c_double),
                ("kernel_ms", ctypes.c_double), ("cells", ctypes.c_int64), ("h2d_bytes", ctypes.c_int64),
                ("d2h_bytes", ctypes.c_int64), ("chunks", ctypes.c_int32), ("launches", ctypes.c_int32),
                ("devices", ctypes.c_int32), ("reserved", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class CudaError(RuntimeError):
    pass


_lib = None

# va_cuda_alloc_fn: char* (*)(size_t bytes, void* user)
_ALLOC_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)


def library_path() -> str:
    return _build.CUDA_PLUGIN


def lib():
    """Load libCUDAKernel.so (building it in-tree first when sources are newer)."""
    global _lib
    if _lib is None:
        path = _build.build_cuda() if os.path.exists(os.path.join(_build.CSRC, "va_kernels.cu")) else _build.CUDA_PLUGIN
        path = os.environ.get("VERSALIGN_CUDA_LIB", path)  # development aid: an alternative build of the same ABI
        if not os.path.exists(path):
            raise CudaError(f"{path} is missing: run `python -m versalignlib_b200.build`")
        L = ctypes.CDLL(path)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.va_cuda_abi_version.restype = ci
        L.va_cuda_last_error.restype = ctypes.c_char_p
        L.va_cuda_device_count.argtypes = [ctypes.POINTER(ci)]
        L.va_cuda_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(ci), ci, ci]
        L.va_cuda_destroy.argtypes = [vp]
        L.va_cuda_destroy.restype = None
        L.va_cuda_set_host_threads.argtypes = [vp, ci]
        L.va_cuda_get_timings.argtypes = [vp, ctypes.POINTER(Timings)]
        sp = ctypes.POINTER(Scoring)
        L.va_cuda_score_ptrs.argtypes = [vp, ci, sp, ci, vp, ci, vp, ci, vp]
        L.va_cuda_align_ptrs.argtypes = [vp, ci, ci, sp, ci, vp, ci, vp, ci, vp, vp, vp, vp]
        L.va_cuda_align_alloc.argtypes = [vp, ci, ci, sp, ci, vp, ci, vp, ci, _ALLOC_FN, vp, vp, vp, vp, vp]
        L.va_cuda_align_records.argtypes = [vp, ci, ci, sp, ci, vp, ci, vp, ci, _ALLOC_FN, vp, vp, ctypes.c_size_t, vp]
        L.va_cuda_score_flat.argtypes = [vp, ci, sp, ci, vp, ci, vp, ci, vp]
        L.va_cuda_align_flat.argtypes = [vp, ci, ci, sp, ci, vp, ci, vp, ci, vp, vp, vp, vp]
        L.va_cuda_score_packed.argtypes = [vp, ci, sp, ci, vp, vp, vp, vp, vp]
        L.va_cuda_align_packed.argtypes = [vp, ci, ci, sp, ci, vp, vp, vp, vp, vp, vp, vp, _ALLOC_FN, vp, vp]
        L.va_cuda_score_device.argtypes = [vp, ci, sp, ci, vp, ci, vp, ci, vp, vp]
        L.va_cuda_align_device.argtypes = [vp, ci, ci, sp, ci, vp, ci, vp, ci, vp, vp, vp, vp, vp]
        L.va_cuda_max_resident_pairs.argtypes = [vp, ci, ci, ci, ctypes.POINTER(ctypes.c_int64)]
        L.va_cuda_int_peak.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_double), vp]
        L.va_cuda_set_profiling.argtypes = [vp, ci]
        L.va_cuda_get_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float * 3)]
        L.va_cuda_plugin_timings.argtypes = [ctypes.POINTER(Timings)]
        _lib = L
    return _lib


def plugin_timings() -> dict | None:
    """Phase timings of the last call made through the CUDAKernel plug-in in this process."""
    t = Timings()
    if lib().va_cuda_plugin_timings(ctypes.byref(t)) != 0:
        return None
    return t.as_dict()


def fasta_load(path: str):
    """va_fasta_load (include/versalign_fasta.h): a FASTA file as (bases uint8[total], offsets int64[n+1],
    max_length) -- the packed layout score_packed / align_packed take.  Host code, no GPU needed."""
    L = lib()
    L.va_fasta_load.argtypes = [ctypes.c_char_p, _ALLOC_FN, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
    blocks = []

    def _alloc(nbytes, _user):
        buf = np.empty(max(int(nbytes), 1), dtype=np.uint8)
        blocks.append(buf)
        return buf.ctypes.data

    cb = _ALLOC_FN(_alloc)
    pb, po = ctypes.c_void_p(), ctypes.c_void_p()
    n, mx = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = L.va_fasta_load(path.encode(), cb, None, ctypes.byref(pb), ctypes.byref(po), ctypes.byref(n), ctypes.byref(mx))
    if rc != 0:
        raise CudaError(f"va_fasta_load({path}) rc={rc}")
    offsets = blocks[1][: (n.value + 1) * 8].view(np.int64).copy()
    bases = blocks[0][: int(offsets[-1])]
    return bases, offsets, int(mx.value)


def device_count() -> int:
    n = ctypes.c_int(0)
    lib().va_cuda_device_count(ctypes.byref(n))
    return n.value


def _u8(a: np.ndarray) -> np.ndarray:
    assert a.dtype == np.uint8 and a.ndim == 2
    return np.ascontiguousarray(a)


def _row_pointers(a: np.ndarray) -> np.ndarray:
    """char* per row of a C-contiguous 2-D array (the reference's char const* const* shape)."""
    return (a.ctypes.data + np.arange(a.shape[0], dtype=np.uint64) * np.uint64(a.strides[0])).astype(np.uint64)


class CudaContext:
    """va_cuda_ctx: streams, pinned staging and device workspace on one or more GPUs."""

    def __init__(self, devices: list[int] | None = None, host_threads: int = 0):
        L = lib()
        self._L = L
        h = ctypes.c_void_p()
        if devices:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = L.va_cuda_create(ctypes.byref(h), arr, len(devices), host_threads)
        else:
            rc = L.va_cuda_create(ctypes.byref(h), None, 0, host_threads)
        if rc != 0:
            raise CudaError(f"va_cuda_create rc={rc}: {L.va_cuda_last_error().decode()}")
        self._h = h

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise CudaError(f"{what} rc={rc}: {self._L.va_cuda_last_error().decode()}")

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.va_cuda_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_host_threads(self, n: int) -> None:
        self._check(self._L.va_cuda_set_host_threads(self._h, n), "va_cuda_set_host_threads")

    def timings(self) -> dict:
        t = Timings()
        self._check(self._L.va_cuda_get_timings(self._h, ctypes.byref(t)), "va_cuda_get_timings")
        return t.as_dict()

    # ---- host buffers ----------------------------------------------------------------
    def score_flat(self, opt: int, reads: np.ndarray, refs: np.ndarray, scoring=(2, -1, -3, -3),
                   out: np.ndarray | None = None) -> np.ndarray:
        reads, refs = _u8(reads), _u8(refs)
        n = reads.shape[0]
        if out is None:
            out = np.zeros(n, dtype=np.int16)
        sc = Scoring(*scoring)
        self._check(self._L.va_cuda_score_flat(self._h, opt, ctypes.byref(sc), n, reads.ctypes.data, reads.shape[1],
                                               refs.ctypes.data, refs.shape[1], out.ctypes.data), "va_cuda_score_flat")
        return out

    def score_ptrs(self, opt: int, reads: np.ndarray, refs: np.ndarray, scoring=(2, -1, -3, -3)) -> np.ndarray:
        reads, refs = _u8(reads), _u8(refs)
        n = reads.shape[0]
        out = np.zeros(n, dtype=np.int16)
        rp, fp = _row_pointers(reads), _row_pointers(refs)
        sc = Scoring(*scoring)
        self._check(self._L.va_cuda_score_ptrs(self._h, opt, ctypes.byref(sc), n, rp.ctypes.data, reads.shape[1],
                                               fp.ctypes.data, refs.shape[1], out.ctypes.data), "va_cuda_score_ptrs")
        return out

    def align_flat(self, opt: int, policy: int, reads: np.ndarray, refs: np.ndarray, scoring=(2, -1, -3, -3)):
        """Returns (aln_read[n,L], aln_ref[n,L], start[n], end_cell[n,2])."""
        reads, refs = _u8(reads), _u8(refs)
        n, L = reads.shape[0], reads.shape[1] + refs.shape[1]
        a = np.zeros((n, L), dtype=np.uint8)
        b = np.zeros((n, L), dtype=np.uint8)
        start = np.zeros(n, dtype=np.int16)
        end = np.zeros((n, 2), dtype=np.int16)
        sc = Scoring(*scoring)
        self._check(self._L.va_cuda_align_flat(self._h, opt, policy, ctypes.byref(sc), n, reads.ctypes.data,
                                               reads.shape[1], refs.ctypes.data, refs.shape[1], a.ctypes.data,
                                               b.ctypes.data, start.ctypes.data, end.ctypes.data), "va_cuda_align_flat")
        return a, b, start, end

    def align_ptrs(self, opt: int, policy: int, reads: np.ndarray, refs: np.ndarray, scoring=(2, -1, -3, -3)):
        reads, refs = _u8(reads), _u8(refs)
        n, L = reads.shape[0], reads.shape[1] + refs.shape[1]
        a = np.zeros((n, L), dtype=np.uint8)
        b = np.zeros((n, L), dtype=np.uint8)
        start = np.zeros(n, dtype=np.int16)
        end = np.zeros((n, 2), dtype=np.int16)
        rp, fp, ap, bp = _row_pointers(reads), _row_pointers(refs), _row_pointers(a), _row_pointers(b)
        sc = Scoring(*scoring)
        self._check(self._L.va_cuda_align_ptrs(self._h, opt, policy, ctypes.byref(sc), n, rp.ctypes.data, reads.shape[1],
                                               fp.ctypes.data, refs.shape[1], ap.ctypes.data, bp.ctypes.data,
                                               start.ctypes.data, end.ctypes.data), "va_cuda_align_ptrs")
        return a, b, start, end

    def align_alloc(self, opt: int, policy: int, reads: np.ndarray, refs: np.ndarray, scoring=(2, -1, -3, -3),
                    fail_after: int | None = None):
        """va_cuda_align_alloc with a Python allocator (blocks are numpy arrays kept alive here).  Returns
        (aln_read[n,L], aln_ref[n,L], start[n], end_cell[n,2]) assembled from the blocks; fail_after = k makes
        the allocator return NULL from its k-th call on (the call must then fail with VA_ERR_MEMORY)."""
        import threading
        reads, refs = _u8(reads), _u8(refs)
        n, L = reads.shape[0], reads.shape[1] + refs.shape[1]
        blocks, lock, calls = {}, threading.Lock(), [0]

        def _alloc(nbytes, _user):
            with lock:
                calls[0] += 1
                if fail_after is not None and calls[0] > fail_after:
                    return None
                buf = np.zeros(max(int(nbytes), 1), dtype=np.uint8)
                blocks[buf.ctypes.data] = buf
                return buf.ctypes.data

        cb = _ALLOC_FN(_alloc)
        pa, pb = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
        start = np.zeros(n, dtype=np.int16)
        end = np.zeros((n, 2), dtype=np.int16)
        rp, fp = _row_pointers(reads), _row_pointers(refs)
        sc = Scoring(*scoring)
        self._check(self._L.va_cuda_align_alloc(self._h, opt, policy, ctypes.byref(sc), n, rp.ctypes.data, reads.shape[1],
                                                fp.ctypes.data, refs.shape[1], cb, None, pa.ctypes.data, pb.ctypes.data,
                                                start.ctypes.data, end.ctypes.data), "va_cuda_align_alloc")
        a = np.stack([blocks[int(x)][:L] for x in pa]) if n else np.zeros((0, L), np.uint8)
        b = np.stack([blocks[int(x)][:L] for x in pb]) if n else np.zeros((0, L), np.uint8)
        return a, b, start, end

    # ---- batch-friendly: offset-addressed sequences, CIGAR out -------------------------
    def score_packed(self, opt: int, reads: np.ndarray, read_off: np.ndarray, refs: np.ndarray, ref_off: np.ndarray,
                     scoring=(2, -1, -3, -3)) -> np.ndarray:
        """reads / refs: 1-D uint8 (all sequences back to back); *_off: int64[n+1]."""
        reads, refs = np.ascontiguousarray(reads, np.uint8), np.ascontiguousarray(refs, np.uint8)
        read_off, ref_off = np.ascontiguousarray(read_off, np.int64), np.ascontiguousarray(ref_off, np.int64)
        n = read_off.shape[0] - 1
        out = np.zeros(n, dtype=np.int16)
        sc = Scoring(*scoring)
        self._check(self._L.va_cuda_score_packed(self._h, opt, ctypes.byref(sc), n, reads.ctypes.data, read_off.ctypes.data,
                                                 refs.ctypes.data, ref_off.ctypes.data, out.ctypes.data), "va_cuda_score_packed")
        return out

    def align_packed(self, opt: int, policy: int, reads: np.ndarray, read_off: np.ndarray, refs: np.ndarray,
                     ref_off: np.ndarray, scoring=(2, -1, -3, -3), out: dict | None = None):
        """Returns (scores[n], coords[n,4] = read_begin, read_end, ref_begin, ref_end, cigar_off[n+1], cigar uint32[]).
        `out`: a dict the call keeps its output arrays in, so a caller that repeats calls of the same size
        reuses them (keys scores, coords, cigar_off, cigar)."""
        reads, refs = np.ascontiguousarray(reads, np.uint8), np.ascontiguousarray(refs, np.uint8)
        read_off, ref_off = np.ascontiguousarray(read_off, np.int64), np.ascontiguousarray(ref_off, np.int64)
        n = read_off.shape[0] - 1
        out = {} if out is None else out

        def _arr(key, shape, dtype):
            a = out.get(key)
            if a is None or a.shape != shape:
                a = out[key] = np.zeros(shape, dtype=dtype)
            return a

        scores, coords, cigar_off = _arr("scores", (n,), np.int16), _arr("coords", (n, 4), np.int32), _arr("cigar_off", (n + 1,), np.int64)

        def _alloc(nbytes, _user):
            words = max(int(nbytes) // 4, 1)
            buf = out.get("cigar")
            if buf is None or buf.shape[0] < words:
                buf = out["cigar"] = np.empty(words + words // 8, dtype=np.uint32)
            return buf.ctypes.data

        cb = _ALLOC_FN(_alloc)
        out_ptr = ctypes.c_void_p()
        sc = Scoring(*scoring)
        self._check(self._L.va_cuda_align_packed(self._h, opt, policy, ctypes.byref(sc), n, reads.ctypes.data,
                                                 read_off.ctypes.data, refs.ctypes.data, ref_off.ctypes.data,
                                                 scores.ctypes.data, coords.ctypes.data, cigar_off.ctypes.data, cb, None,
                                                 ctypes.byref(out_ptr)), "va_cuda_align_packed")
        cigar = out["cigar"][: int(cigar_off[n])] if "cigar" in out else np.zeros(0, np.uint32)
        return scores, coords, cigar_off, cigar

    # ---- device-resident (torch tensors supply the memory and the stream) --------------
    def score_device(self, opt: int, d_reads, d_refs, d_scores, scoring=(2, -1, -3, -3), stream: int = 0) -> None:
        n = d_reads.shape[0]
        sc = Scoring(*scoring)
        self._check(self._L.va_cuda_score_device(self._h, opt, ctypes.byref(sc), n, d_reads.data_ptr(), d_reads.shape[1],
                                                 d_refs.data_ptr(), d_refs.shape[1], d_scores.data_ptr(), stream),
                    "va_cuda_score_device")

    def align_device(self, opt: int, policy: int, d_reads, d_refs, d_aln_read, d_aln_ref, d_start, d_end_cell=None,
                     scoring=(2, -1, -3, -3), stream: int = 0) -> None:
        n = d_reads.shape[0]
        sc = Scoring(*scoring)
        self._check(self._L.va_cuda_align_device(self._h, opt, policy, ctypes.byref(sc), n, d_reads.data_ptr(),
                                                 d_reads.shape[1], d_refs.data_ptr(), d_refs.shape[1],
                                                 d_aln_read.data_ptr(), d_aln_ref.data_ptr(), d_start.data_ptr(),
                                                 d_end_cell.data_ptr() if d_end_cell is not None else None, stream),
                    "va_cuda_align_device")

    def max_resident_pairs(self, align: bool, read_length: int, ref_length: int) -> int:
        out = ctypes.c_int64(0)
        self._check(self._L.va_cuda_max_resident_pairs(self._h, 1 if align else 0, read_length, ref_length,
                                                       ctypes.byref(out)), "va_cuda_max_resident_pairs")
        return out.value

    def set_profiling(self, on: bool) -> None:
        self._check(self._L.va_cuda_set_profiling(self._h, 1 if on else 0), "va_cuda_set_profiling")

    def kernel_ms(self) -> tuple[float, float, float]:
        """(prep, fill, traceback) milliseconds of the last device-resident call (profiling on)."""
        ms = (ctypes.c_float * 3)()
        self._check(self._L.va_cuda_get_kernel_ms(self._h, ctypes.byref(ms)), "va_cuda_get_kernel_ms")
        return float(ms[0]), float(ms[1]), float(ms[2])

    def int_peak(self, kind: int, stream: int = 0) -> float:
        """Measured integer-pipe throughput in lane-ops/s (kind: 0 s32, 1 s16x2, 2 s16x2.relu, 3 vimax3 s16x2)."""
        out = ctypes.c_double(0)
        self._check(self._L.va_cuda_int_peak(self._h, kind, ctypes.byref(out), stream), "va_cuda_int_peak")
        return out.value
