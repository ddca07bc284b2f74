"""Randomised parity stress (run by hand on a GPU box; not collected by pytest): random batch shapes, length
mixes, dirt, scorings, modes and policies on ONE context -- so every call reuses the workspace the previous
ones left behind -- each compared bit-exactly with the CPU oracle.
usage: python tests/stress_random.py [seconds] [seed]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import binding as ora  # noqa: E402
from tests.helpers import used_region_equal  # noqa: E402
from versalignlib_b200 import capi, synth  # noqa: E402


def random_case(rng):
    kind = rng.choice(["uniform", "mixed", "mixed_reads", "dirty", "long", "many"], p=[0.28, 0.28, 0.14, 0.14, 0.12, 0.04])
    if kind == "long":  # intra-task kernels, global traceback queues
        n = int(rng.integers(1, 24))
        rl, fl = int(rng.integers(300, 2600)), int(rng.integers(300, 3200))
    elif kind == "many":  # several chunks through the host pipeline
        n = int(rng.integers(70_000, 150_000))
        rl, fl = int(rng.integers(8, 48)), int(rng.integers(8, 48))
    else:
        n = int(rng.choice([1, 2, 3, 31, 64, 65, 127, 500, 2000, 9000]))
        rl, fl = int(rng.integers(1, 260)), int(rng.integers(1, 260))
    seed = int(rng.integers(1 << 30))
    if kind == "long" and rng.random() < 0.4:  # long pairs of mixed lengths (partial last passes, duos of unequal reads)
        reads, refs, _, _ = synth.mixed_batch(n, int(rng.integers(1, 300)), max(rl, fl, 1100), p_sub=0.1, q_indel=0.02, seed=seed)
    elif kind in ("uniform", "long", "many"):
        reads, refs = synth.uniform_batch(n, rl, fl, p_sub=float(rng.choice([0.05, 0.3, 0.75])), q_indel=float(rng.choice([0, 0.03])), seed=seed)
    else:
        lo = max(1, min(rl, fl) // 3)
        reads, refs, _, _ = synth.mixed_batch(n, lo, max(rl, fl, lo), p_sub=0.1, q_indel=0.02, seed=seed)
        if kind == "mixed_reads":  # refs of one length, reads of every length: end-aligned NW duos
            refs = synth.random_seqs(np.random.Generator(np.random.PCG64(seed)), n, refs.shape[1])
        if kind == "dirty":
            reads, refs = synth.sprinkle(seed, reads, 0.03), synth.sprinkle(seed + 1, refs, 0.02)
    sc = (int(rng.choice([1, 2, 5, 20])), int(rng.choice([0, -1, -4, -30])), int(rng.choice([-1, -3, -7, -50, 0])), int(rng.choice([-1, -3, -7, -40])))
    # stay inside the exactness domain: no cell may leave int16
    if sc[0] * min(reads.shape[1], refs.shape[1]) > 30000 or max(abs(sc[2]), abs(sc[3])) * (reads.shape[1] + refs.shape[1] + 4) > 30000:
        sc = (2, -1, -3, -3)
    return kind, reads, refs, sc


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
    t0, cases, failures = time.time(), 0, []
    with capi.CudaContext(devices=[0]) as ctx:
        while time.time() - t0 < budget:
            kind, reads, refs, sc = random_case(rng)
            tag = (kind, reads.shape, refs.shape, sc)
            for opt in (ora.SW, ora.NW):
                if not np.array_equal(ctx.score_flat(opt, reads, refs, sc), ora.score(opt, reads, refs, sc)):
                    failures.append(("score", opt) + tag)
                policy = int(rng.integers(2))
                a, b, start, end = ctx.align_flat(opt, policy, reads, refs, sc)
                oa, ob, ostart, oend = ora.align(opt, policy, reads, refs, sc)
                if not (np.array_equal(start, ostart) and np.array_equal(end, oend) and used_region_equal(a, b, start, oa, ob, ostart).size == 0):
                    failures.append(("align", opt, policy) + tag)
            # the packed / CIGAR entry points on the same batch (sequences trimmed of their '\0' padding; the
            # semantics are the reference's on the batch padded to ITS maximum lengths)
            # (not the dirty decks: their junk alphabet contains '-', which the string -> CIGAR derivation of the
            # expected result cannot tell from a gap)
            if kind != "dirty" and reads.shape[0] <= 2000 and rng.random() < 0.5:
                pr, ro = synth.pack_batch(reads)
                pf, fo = synth.pack_batch(refs)
                tr = np.ascontiguousarray(reads[:, :max(int(np.diff(ro).max()), 1)])
                tf = np.ascontiguousarray(refs[:, :max(int(np.diff(fo).max()), 1)])
                # interior '\0' bytes would change the trimmed lengths' meaning: only clean tails
                if (np.diff(ro) == (tr != 0).sum(axis=1)).all() and (np.diff(fo) == (tf != 0).sum(axis=1)).all() and np.diff(ro).min() > 0 and np.diff(fo).min() > 0:
                    for opt in (ora.SW, ora.NW):
                        if not np.array_equal(ctx.score_packed(opt, pr, ro, pf, fo, sc), ora.score(opt, tr, tf, sc)):
                            failures.append(("score_packed", opt) + tag)
                        policy = int(rng.integers(2))
                        scores, coords, coff, cigar = ctx.align_packed(opt, policy, pr, ro, pf, fo, sc)
                        want = synth.cigar_from_strings(*ora.align(opt, policy, tr, tf, sc))
                        if not (np.array_equal(coords, want[0]) and np.array_equal(coff, want[1]) and np.array_equal(cigar, want[2])):
                            failures.append(("align_packed", opt, policy) + tag)
            # the affine-gap variant (general kernel) against its own oracle; gap_open = 0 must also equal the linear result
            if kind not in ("long", "many") and sc[2] <= 0 and sc[3] <= 0 and rng.random() < 0.3:
                gap_open = int(rng.choice([0, -2, -9]))
                for opt in (ora.SW, ora.NW):
                    o = capi.affine_opt(opt, gap_open)
                    if not np.array_equal(ctx.score_flat(o, reads, refs, sc), ora.score_affine(opt, reads, refs, sc, gap_open)):
                        failures.append(("score_affine", opt, gap_open) + tag)
                    a, b, start, end = ctx.align_flat(o, 0, reads, refs, sc)
                    oa, ob, ostart, oend, _ = ora.align_affine(opt, reads, refs, sc, gap_open)
                    if not (np.array_equal(start, ostart) and np.array_equal(end, oend) and used_region_equal(a, b, start, oa, ob, ostart).size == 0):
                        failures.append(("align_affine", opt, gap_open) + tag)
            cases += 1
            if failures:
                break
    print(f"{cases} random cases x 4 functions in {time.time() - t0:.0f} s: " + ("all bit-exact" if not failures else f"FAILED {failures[:3]}"))
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
