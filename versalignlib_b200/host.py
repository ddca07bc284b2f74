"""Python face of the driver-side loader (csrc/plugin_host.cpp): loads an AlignmentKernel
plug-in -- ours or one of the reference's -- the way the reference driver does
(versalignUtil.cpp:45-76 DLL_init, main.cpp:227-238 get_kernel) and calls its two virtual
methods on host buffers.  Mirrors the reference's calling convention: parameters by key,
batch-wide read_length / ref_length, '\\0' padded fixed-length sequences.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import build as _build

SW, NW = 0, 1

_lib = None


def _host_lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_build.build_host())
        L.vah_create.restype = ctypes.c_void_p
        L.vah_set_param.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        L.vah_unset_param.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.vah_set_verbosity.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.vah_error.restype = ctypes.c_char_p
        L.vah_error.argtypes = [ctypes.c_void_p]
        L.vah_log_count.restype = ctypes.c_long
        L.vah_log_count.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.vah_load.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.vah_respawn.argtypes = [ctypes.c_void_p]
        L.vah_stage.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                ctypes.c_int, ctypes.c_int]
        L.vah_last_call_seconds.restype = ctypes.c_double
        L.vah_last_call_seconds.argtypes = [ctypes.c_void_p]
        L.vah_score.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.vah_align.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.vah_fetch_alignments.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.vah_alignments_terminated.argtypes = [ctypes.c_void_p]
        L.vah_drop_alignments.argtypes = [ctypes.c_void_p]
        L.vah_close.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


class PluginError(RuntimeError):
    pass


class PluginHost:
    """One loaded kernel library + one spawned AlignmentKernel instance.

    scoring = (score_match, score_mismatch, score_gap_read, score_gap_ref), the keys of
    CustomParameters.h:9-47; extra maps additional keys (e.g. cuda_traceback_policy).
    """

    def __init__(self, library: str, read_length: int, ref_length: int, scoring=(2, -1, -3, -3),
                 num_threads: int = 1, extra: dict | None = None, verbosity: int = 1, omit: tuple = ()):
        L = _host_lib()
        self._L = L
        self._h = ctypes.c_void_p(L.vah_create())
        self.read_length, self.ref_length = read_length, ref_length
        params = {
            "score_match": scoring[0], "score_mismatch": scoring[1], "score_gap_read": scoring[2],
            "score_gap_ref": scoring[3], "read_length": read_length, "ref_length": ref_length,
            "num_threads": num_threads,
        }
        params.update(extra or {})
        for k, v in params.items():
            if k not in omit:
                L.vah_set_param(self._h, k.encode(), int(v))
        L.vah_set_verbosity(self._h, verbosity)
        if L.vah_load(self._h, library.encode()) != 0:
            msg = L.vah_error(self._h).decode()
            L.vah_close(self._h)
            self._h = None
            raise PluginError(msg)
        self._n = 0

    def set_param(self, key: str, value: int, respawn: bool = True) -> None:
        self._L.vah_set_param(self._h, key.encode(), int(value))
        if respawn and self._L.vah_respawn(self._h) != 0:
            raise PluginError(self._L.vah_error(self._h).decode())

    def stage(self, reads: np.ndarray, refs: np.ndarray, scattered: bool = True) -> None:
        assert reads.dtype == np.uint8 and refs.dtype == np.uint8
        assert reads.shape == (reads.shape[0], self.read_length) and refs.shape == (reads.shape[0], self.ref_length)
        reads, refs = np.ascontiguousarray(reads), np.ascontiguousarray(refs)
        self._n = reads.shape[0]
        self._L.vah_stage(self._h, self._n, reads.ctypes.data, self.read_length, refs.ctypes.data, self.ref_length,
                          1 if scattered else 0)

    def score_staged(self, opt: int, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.zeros(self._n, dtype=np.int16)
        if self._L.vah_score(self._h, opt, out.ctypes.data) != 0:
            raise PluginError(self._L.vah_error(self._h).decode())
        return out

    def align_staged(self, opt: int, fetch: bool = True):
        if self._L.vah_align(self._h, opt) != 0:
            raise PluginError(self._L.vah_error(self._h).decode())
        if not fetch:
            return None
        return self.fetch_alignments()

    def fetch_alignments(self):
        n, L = self._n, self.read_length + self.ref_length
        a = np.zeros((n, L), dtype=np.uint8)
        b = np.zeros((n, L), dtype=np.uint8)
        f = np.zeros((n, 4), dtype=np.int16)
        if self._L.vah_fetch_alignments(self._h, a.ctypes.data, b.ctypes.data, f.ctypes.data) != 0:
            raise PluginError(self._L.vah_error(self._h).decode())
        return a, b, f

    def alignments_terminated(self) -> bool:
        return self._L.vah_alignments_terminated(self._h) == 1

    def drop_alignments(self) -> None:
        self._L.vah_drop_alignments(self._h)

    @property
    def last_call_seconds(self) -> float:
        return self._L.vah_last_call_seconds(self._h)

    def log_count(self, bucket: int) -> int:
        return self._L.vah_log_count(self._h, bucket)

    # convenience: one-shot calls
    def score_alignments(self, opt: int, reads: np.ndarray, refs: np.ndarray, scattered: bool = False) -> np.ndarray:
        self.stage(reads, refs, scattered)
        return self.score_staged(opt)

    def compute_alignments(self, opt: int, reads: np.ndarray, refs: np.ndarray, scattered: bool = False):
        self.stage(reads, refs, scattered)
        return self.align_staged(opt)

    def close(self) -> None:
        if self._h is not None:
            self._L.vah_close(self._h)
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
