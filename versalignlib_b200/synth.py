"""Seeded synthetic read/ref batches in the reference driver's input convention.

The reference ships no test set (its driver reads ../testset/*.fa, main.cpp:86-90, which is
not in the tree), so every parity and bench input comes from here.  A batch is two uint8
arrays of shape (n, read_length) and (n, ref_length): each row is one sequence, '\\0'
padded to the batch-wide maximum and NOT NUL terminated -- exactly what the reference's
pad() (versalignUtil.cpp:17-33) hands to the kernels.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

# seed = 0x5EED0000 + config number
BASE_SEED = 0x5EED0000


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def random_seqs(rng: np.random.Generator, n: int, length: int) -> np.ndarray:
    return ACGT[rng.integers(0, 4, size=(n, length), dtype=np.uint8)]


def mutate_from_ref(rng, refs: np.ndarray, read_len: int, p_sub: float, q_indel: float) -> np.ndarray:
    """read = window of ref at a uniform offset, with per-base substitutions (p_sub) and
    insertions/deletions (q_indel, half each)."""
    n, ref_len = refs.shape
    slack = max(ref_len - read_len, 0)
    offset = rng.integers(0, slack + 1, size=(n, 1))
    if q_indel > 0:
        u = rng.random((n, read_len))
        step = np.ones((n, read_len), dtype=np.int64)
        step[u < q_indel / 2] = 0            # insertion: emit a random base, do not advance
        step[(u >= q_indel / 2) & (u < q_indel)] = 2  # deletion: skip one ref base
        src = offset + np.cumsum(step, axis=1) - 1
        inserted = step == 0
    else:
        src = offset + np.arange(read_len)[None, :]
        inserted = None
    src = np.clip(src, 0, ref_len - 1)
    reads = np.take_along_axis(refs, src, axis=1)
    if inserted is not None:
        reads = np.where(inserted, random_seqs(rng, n, read_len), reads)
    if p_sub > 0:
        sub = rng.random((n, read_len)) < p_sub
        # substitute with one of the three OTHER bases
        code = np.searchsorted(ACGT, reads)  # ACGT is sorted: A<C<G<T
        new = ACGT[(code + rng.integers(1, 4, size=(n, read_len))) % 4]
        reads = np.where(sub, new, reads)
    return np.ascontiguousarray(reads.astype(np.uint8))


def uniform_batch(n: int, read_len: int, ref_len: int, p_sub: float = 0.1, q_indel: float = 0.0,
                  seed: int = BASE_SEED, independent: bool = False):
    """All pairs have the full length (configs C1, C2, C4, C5)."""
    rng = _rng(seed)
    refs = random_seqs(rng, n, ref_len)
    if independent:
        reads = random_seqs(rng, n, read_len)
    else:
        reads = mutate_from_ref(rng, refs, read_len, p_sub, q_indel)
    return reads, refs


def mixed_batch(n: int, min_len: int, max_len: int, p_sub: float = 0.1, q_indel: float = 0.0,
                seed: int = BASE_SEED + 3):
    """Config C3: read length ~U{min..max}, ref length ~U{read..max}; buffers padded with
    '\\0' to max_len.  Returns (reads, refs, read_lens, ref_lens)."""
    rng = _rng(seed)
    refs = random_seqs(rng, n, max_len)
    reads = mutate_from_ref(rng, refs, max_len, p_sub, q_indel)
    rl = rng.integers(min_len, max_len + 1, size=n)
    fl = rng.integers(rl, max_len + 1)
    col = np.arange(max_len)[None, :]
    reads = np.where(col < rl[:, None], reads, 0).astype(np.uint8)
    refs = np.where(col < fl[:, None], refs, 0).astype(np.uint8)
    return np.ascontiguousarray(reads), np.ascontiguousarray(refs), rl.astype(np.int32), fl.astype(np.int32)


def sprinkle(rng_seed: int, seqs: np.ndarray, frac: float, alphabet: bytes = b"NnacgtX-") -> np.ndarray:
    """Replace a fraction of the non-pad bytes by N / lower case / junk (edge-case decks)."""
    rng = _rng(rng_seed)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    hit = (rng.random(seqs.shape) < frac) & (seqs != 0)
    repl = alpha[rng.integers(0, len(alpha), size=seqs.shape)]
    return np.where(hit, repl, seqs).astype(np.uint8)


def edge_deck(read_len: int, ref_len: int):
    """Hand-made corner cases (SURVEY.md 8(d) "edge deck"), padded to the given lengths."""
    def row(s: bytes, L: int) -> np.ndarray:
        s = s[:L]
        return np.frombuffer(s + b"\0" * (L - len(s)), dtype=np.uint8)

    cases = [
        (b"ACGTACGTAC", b"ACGTACGTAC"),              # identical
        (b"A", b"A"),                                # length 1
        (b"A", b"C"),
        (b"AAAAAAAA", b"CCCCCCCC"),                  # no match at all (score 0)
        (b"NNNNNNNN", b"NNNNNNNN"),                  # all N
        (b"acgtacgt", b"ACGTACGT"),                  # lower case
        (b"ACGTNACGT", b"ACGTAACGT"),                # embedded N in read
        (b"ACGTAACGT", b"ACGNTAACGT"),               # embedded N in ref
        (b"ACGT-ACGT", b"ACGTACGT"),                 # junk byte inside the read
        (b"ACGTACGT", b"ACG\0ACGT"),                 # embedded NUL inside the ref
        (b"", b"ACGT"),                              # empty read
        (b"ACGT", b""),                              # empty ref
        (b"", b""),
        (b"NACGT", b"ACGT"),                         # invalid first char
        (b"ACGT" * 64, b"ACGT" * 64),                # full length repeats
        (b"ACGT" * 64, b"TGCA" * 64),
        (b"GATTACA", b"GCATGCU"),
        (b"ACGTACGTTT", b"TTACGTACGT"),
        (bytes([0xC1, 0x41, 0x43, 0xE7, 0x47]), b"AACGG"),  # bytes >= 0x80
    ]
    reads = np.stack([row(a, read_len) for a, _ in cases])
    refs = np.stack([row(b, ref_len) for _, b in cases])
    return np.ascontiguousarray(reads), np.ascontiguousarray(refs)


def pack_batch(seqs: np.ndarray, lens: np.ndarray | None = None):
    """(n, L) padded rows -> (flat uint8, int64 offsets[n+1]); lens defaults to the '\\0'-trimmed lengths."""
    n, L = seqs.shape
    if lens is None:
        nz = seqs != 0
        lens = np.where(nz.any(axis=1), L - np.argmax(nz[:, ::-1], axis=1), 0)
    lens = np.asarray(lens, dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    mask = np.arange(L)[None, :] < lens[:, None]
    return np.ascontiguousarray(seqs[mask]), off


def cigar_from_strings(aln_read: np.ndarray, aln_ref: np.ndarray, start: np.ndarray, end_cell: np.ndarray):
    """What the packed entry points return, derived from the reference-style outputs (gapped strings,
    right-aligned, used region [start, L-1)): coords[n,4], cigar_off[n+1], cigar (BAM encoding)."""
    n, L = aln_read.shape
    coords = np.zeros((n, 4), dtype=np.int32)
    off = np.zeros(n + 1, dtype=np.int64)
    runs = []
    dash = ord("-")
    for i in range(n):
        s0 = max(int(start[i]), 0)
        a, b = aln_read[i, s0:L - 1], aln_ref[i, s0:L - 1]
        op = np.where(b == dash, 1, np.where(a == dash, 2, 0)).astype(np.int64)
        used_read, used_ref = int((op != 2).sum()), int((op != 1).sum())
        coords[i] = (end_cell[i, 0] + 1 - used_read, end_cell[i, 0] + 1, end_cell[i, 1] + 1 - used_ref, end_cell[i, 1] + 1)
        if op.size:
            cut = np.flatnonzero(np.diff(op)) + 1
            starts = np.concatenate(([0], cut))
            lens = np.diff(np.concatenate((starts, [op.size])))
            runs.append(((lens << 4) | op[starts]).astype(np.uint32))
            off[i + 1] = off[i] + starts.size
        else:
            off[i + 1] = off[i]
    cigar = np.concatenate(runs) if runs else np.zeros(0, np.uint32)
    return coords, off, cigar


def mixed_batch_torch(n: int, min_len: int, max_len: int, p_sub: float, seed: int, device, slice_pairs: int = 1 << 20):
    """Config C3 at full size, generated on the GPU (the numpy generator above needs minutes for 10 M pairs):
    same recipe -- ref uniform ACGT, read = ref with per-base substitutions, read length ~U{min..max}, ref length
    ~U{read..max}, '\\0' padding to max_len.  Returns device tensors (reads[n,max_len], refs[n,max_len] uint8,
    read_lens[n], ref_lens[n] int32).  Bench input only: parity tests use the numpy generators."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    reads = torch.empty((n, max_len), dtype=torch.uint8, device=device)
    refs = torch.empty((n, max_len), dtype=torch.uint8, device=device)
    rl = torch.randint(min_len, max_len + 1, (n,), generator=g, device=device, dtype=torch.int32)
    fl = (rl + (torch.rand((n,), generator=g, device=device) * (max_len + 1 - rl).float()).int()).clamp_(max=max_len)
    col = torch.arange(max_len, device=device, dtype=torch.int32)[None, :]
    for lo in range(0, n, slice_pairs):
        hi = min(n, lo + slice_pairs)
        code = torch.randint(0, 4, (hi - lo, max_len), generator=g, device=device, dtype=torch.uint8)
        sub = torch.rand((hi - lo, max_len), generator=g, device=device) < p_sub
        shift = torch.randint(1, 4, (hi - lo, max_len), generator=g, device=device, dtype=torch.uint8)
        rcode = torch.where(sub, (code + shift) % 4, code)
        refs[lo:hi] = torch.where(col < fl[lo:hi, None], acgt[code.long()], torch.zeros((), dtype=torch.uint8, device=device))
        reads[lo:hi] = torch.where(col < rl[lo:hi, None], acgt[rcode.long()], torch.zeros((), dtype=torch.uint8, device=device))
        del code, sub, shift, rcode
    return reads, refs, rl, fl


def pack_batch_torch(seqs, lens, slice_pairs: int = 1 << 20):
    """Device (n, L) padded rows + lengths -> page-locked host arrays (flat uint8, int64 offsets[n+1]) as numpy views."""
    import torch
    n, L = seqs.shape
    off = torch.zeros(n + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(lens.to(torch.int64).cpu(), 0)
    flat = torch.empty(int(off[-1]), dtype=torch.uint8).pin_memory()
    col = torch.arange(L, device=seqs.device, dtype=torch.int32)[None, :]
    for lo in range(0, n, slice_pairs):
        hi = min(n, lo + slice_pairs)
        part = seqs[lo:hi][col < lens[lo:hi, None]]
        flat[int(off[lo]):int(off[hi])].copy_(part)
    return flat.numpy(), off.pin_memory().numpy()
