#!/usr/bin/env python
"""Generate tests/golden/*.npz: small seeded inputs together with the outputs of the REFERENCE'S
OWN kernels, executed (libDefaultKernel.so / libSSEKernel.so / libAVXKernel.so compiled from
/root/reference by oracle/Makefile and loaded through the reference's plug-in boundary).

The reference ships no golden vectors of its own, and /root/reference does
not exist on the GPU box, so these fixtures are how the reference's behaviour travels: the oracle
and the CUDA path are both checked against them.

Run here (needs oracle/_ref):   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import binding as ora  # noqa: E402
from versalignlib_b200 import synth  # noqa: E402
from versalignlib_b200.host import PluginHost  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PARAMS = [(2, -1, -3, -3), (3, -2, -1, -4), (5, -4, -1, -7)]


def decks():
    out = {}
    out["c1_100x150"] = synth.uniform_batch(64, 100, 150, p_sub=0.10, seed=synth.BASE_SEED + 1)
    out["c2_150x150"] = synth.uniform_batch(64, 150, 150, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 2)
    out["c5_random64x96"] = synth.uniform_batch(64, 64, 96, independent=True, seed=synth.BASE_SEED + 5)
    r, f, _, _ = synth.mixed_batch(64, 30, 110, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 3)
    out["mixed30-110"] = (r, f)
    out["dirty"] = (synth.sprinkle(21, r, 0.03, b"Nnacgt"), synth.sprinkle(22, f, 0.03, b"Nnacgt"))
    er, ef = synth.edge_deck(48, 64)
    # the Default kernel indexes a table with a signed char: keep bytes < 0x80 in the shared deck
    er, ef = np.where(er >= 0x80, ord("X"), er).astype(np.uint8), np.where(ef >= 0x80, ord("X"), ef).astype(np.uint8)
    reps = 32 // len(er) + 1
    out["edge"] = (np.ascontiguousarray(np.tile(er, (reps, 1))[:32]), np.ascontiguousarray(np.tile(ef, (reps, 1))[:32]))
    return out


def run_reference(kernel, reads, refs, sc):
    lib = ora.ref_lib(kernel)
    assert lib, f"oracle/_ref/lib{kernel}Kernel.so missing: run make -C oracle"
    res = {}
    with PluginHost(lib, reads.shape[1], refs.shape[1], sc, num_threads=1, verbosity=0) as h:
        h.stage(reads, refs, scattered=False)
        for opt, name in ((0, "sw"), (1, "nw")):
            res[f"{name}_score"] = h.score_staged(opt)
            a, b, f = h.align_staged(opt)
            L = reads.shape[1] + refs.shape[1]
            # keep only what the reference defines: bytes [start, L-1) of both strings + start
            col = np.arange(L)[None, :]
            used = (col >= f[:, 0:1]) & (col < L - 1)
            res[f"{name}_aln_read"] = np.where(used, a, 0).astype(np.uint8)
            res[f"{name}_aln_ref"] = np.where(used, b, 0).astype(np.uint8)
            res[f"{name}_start"] = f[:, 0].copy()
    return res


def main():
    for label, (reads, refs) in decks().items():
        blob = {"reads": reads, "refs": refs, "params": np.array(PARAMS, dtype=np.int32)}
        for pi, sc in enumerate(PARAMS):
            sse = run_reference("SSE", reads, refs, sc)
            avx = run_reference("AVX", reads, refs, sc)
            dfl = run_reference("Default", reads, refs, sc)
            for k in sse:
                assert np.array_equal(sse[k], avx[k]), (label, sc, k, "SSE vs AVX")
            # scores: the full short as SSE/AVX store it (Default writes one byte only)
            for m in ("sw", "nw"):
                assert np.array_equal(dfl[f"{m}_score"].view(np.uint8)[0::2], sse[f"{m}_score"].view(np.uint8)[0::2])
                blob[f"p{pi}_{m}_score"] = sse[f"{m}_score"]
                for part in ("aln_read", "aln_ref", "start"):
                    blob[f"p{pi}_{m}_{part}_default"] = dfl[f"{m}_{part}"]
                    blob[f"p{pi}_{m}_{part}_simd"] = sse[f"{m}_{part}"]
        path = os.path.join(HERE, f"{label}.npz")
        np.savez_compressed(path, **blob)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
