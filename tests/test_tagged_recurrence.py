"""The algebra of the tagged NW align kernel (versalignlib_b200/csrc/va_nw.cu), checked on the CPU.

The kernel carries 4V + tag in 16-bit lanes, V the matrix in shifted coordinates (V = H - gap_ref*I - gap_read*J), and
computes   t = max(first + 1, second),  h = max(diag + 4s' + 2, t),  x = h & ~3   with s' = s - gap_ref - gap_read and
first = UP (Default/OpenCL policy) or LEFT (SSE/AVX policy).  Claims: (a) h & 3 is the pointer the reference's rule
(DefaultKernel.cpp:338-346 / SSEKernel.cpp:646-659: DIAG before the tie winner before the other) gives on the unshifted
matrix, in every cell; (b) x is 4x the shifted value; (c) walking those tags back from the oracle's end cell reproduces
the oracle's alignment -- which tests/test_oracle_vs_reference.py pins against the reference's own kernels.
This is a model of the arithmetic in Python integers wrapped to 16 bits, not the kernel: the GPU parity tests cover that."""
import numpy as np
import pytest

from oracle import binding as ora
from versalignlib_b200 import synth

DIAG, FIRST, SECOND = 2, 1, 0


def s16(v: int) -> int:
    v &= 0xFFFF
    return v - 0x10000 if v & 0x8000 else v


def tagged_fill(read: bytes, ref: bytes, sc, policy: int):
    match, mismatch, g_read, g_ref = sc
    m, n = len(read), len(ref)
    off = g_ref + g_read
    W = [[0] * (n + 1) for _ in range(m + 1)]    # clean values x = 4V
    tags = [[None] * (n + 1) for _ in range(m + 1)]
    for J in range(n + 1):
        W[0][J] = s16(4 * (-g_read) * J)          # H(0,J) = 0  ->  V = -gap_read*J
    for I in range(1, m + 1):
        W[I][0] = 0                               # H(I,0) = I*gap_ref  ->  V = 0
        for J in range(1, n + 1):
            s = match if read[I - 1] == ref[J - 1] else mismatch
            entry = 4 * (s - off) + 2
            assert -128 <= entry <= 127
            up, left, diag = W[I - 1][J], W[I][J - 1], W[I - 1][J - 1]
            t = max(s16(up + 1), left) if policy == 0 else max(s16(left + 1), up)
            h = max(s16(diag + entry), t)
            tags[I][J] = h & 3
            W[I][J] = s16(h & ~3)
    return W, tags


def reference_rule(read: bytes, ref: bytes, sc, policy: int):
    match, mismatch, g_read, g_ref = sc
    m, n = len(read), len(ref)
    H = [[0] * (n + 1) for _ in range(m + 1)]
    ptr = [[None] * (n + 1) for _ in range(m + 1)]
    for I in range(1, m + 1):
        H[I][0] = I * g_ref
        for J in range(1, n + 1):
            d = H[I - 1][J - 1] + (match if read[I - 1] == ref[J - 1] else mismatch)
            u, l = H[I - 1][J] + g_ref, H[I][J - 1] + g_read
            if d >= u and d >= l:
                ptr[I][J], H[I][J] = "D", d
            elif (u >= l) if policy == 0 else not (l >= u):
                ptr[I][J], H[I][J] = "U", u
            else:
                ptr[I][J], H[I][J] = "L", l
    return H, ptr


def moves_of_oracle(a: np.ndarray, b: np.ndarray, start: int):
    out = []
    for x, y in zip(a[start:-1].tobytes(), b[start:-1].tobytes()):
        out.append("L" if x == ord("-") else "U" if y == ord("-") else "D")
    return out[::-1]  # last column first: the order of the walk


@pytest.mark.parametrize("sc", [(2, -1, -3, -3), (3, -2, -1, -4), (1, 0, -7, -1), (5, -4, -1, -7), (2, -9, -2, -2), (2, -1, 0, -2)])
@pytest.mark.parametrize("policy", [0, 1])
def test_tags_are_the_reference_pointers(sc, policy):
    reads, refs = synth.uniform_batch(24, 37, 45, p_sub=0.25, q_indel=0.08, seed=91 + policy)
    first, second = ("U", "L") if policy == 0 else ("L", "U")
    name = {DIAG: "D", FIRST: first, SECOND: second}
    oa, ob, ostart, oend = ora.align(ora.NW, policy, reads, refs, sc)
    for p in range(reads.shape[0]):
        rd, rf = reads[p].tobytes(), refs[p].tobytes()
        W, tags = tagged_fill(rd, rf, sc, policy)
        H, ptr = reference_rule(rd, rf, sc, policy)
        m, n = len(rd), len(rf)
        for I in range(1, m + 1):
            for J in range(1, n + 1):
                assert name[tags[I][J]] == ptr[I][J], (p, I, J)
                assert W[I][J] == 4 * (H[I][J] - sc[3] * I - sc[2] * J), (p, I, J)  # (b): x = 4V
        # (c) the walk from the oracle's end cell
        i, j = int(oend[p][0]), int(oend[p][1])
        walk = []
        while i >= 0 and j >= 0:
            mv = name[tags[i + 1][j + 1]]
            walk.append(mv)
            i -= mv != "L"
            j -= mv != "U"
        walk += ["U"] * (i + 1) if j < 0 else []  # matrix column 0: up to row 0 (DefaultKernel.cpp:304)
        assert walk == moves_of_oracle(oa[p], ob[p], int(ostart[p])), p
