"""Shared comparison helpers for the parity tests."""
import numpy as np


def used_region_equal(a_read, a_ref, a_start, b_read, b_ref, b_start):
    """Parity rule for alignments: compare `start` and bytes
    [start, aln_length-1) of both gapped strings; bytes before start are undefined in
    the reference.  Returns indices of differing pairs."""
    n, L = a_read.shape
    a_start = np.asarray(a_start).astype(np.int64)
    b_start = np.asarray(b_start).astype(np.int64)
    bad = a_start != b_start
    col = np.arange(L)[None, :]
    mask = (col >= np.clip(a_start, 0, L)[:, None]) & (col < L - 1)
    bad |= ((a_read != b_read) & mask).any(axis=1)
    bad |= ((a_ref != b_ref) & mask).any(axis=1)
    return np.nonzero(bad)[0]
