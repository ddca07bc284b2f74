"""Small all-modes deck for compute-sanitizer (memcheck / racecheck): every kernel -- inter-task packed, intra-task
packed (the 700 x 1100 deck: CTA-pipelined wavefront passes, warp-per-pair traceback), general -- both pointer
policies, ragged sizes, all result containers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from versalignlib_b200 import capi, synth  # noqa: E402


def main():
    decks = [synth.uniform_batch(130, 100, 150, p_sub=0.1, q_indel=0.02, seed=1),
             synth.uniform_batch(67, 33, 47, p_sub=0.2, seed=2),
             synth.mixed_batch(97, 30, 130, p_sub=0.1, q_indel=0.02, seed=3)[:2],
             synth.edge_deck(48, 64),
             synth.uniform_batch(10, 700, 1100, p_sub=0.1, seed=4)]
    dirty = (synth.sprinkle(5, decks[2][0], 0.03), synth.sprinkle(6, decks[2][1], 0.03))
    decks.append(dirty)
    with capi.CudaContext(devices=[0]) as ctx:
        for reads, refs in decks:
            for sc in [(2, -1, -3, -3), (3, -2, -1, -4), (200, -150, -300, -300)]:
                for opt in (0, 1):
                    ctx.score_flat(opt, reads, refs, sc)
                    for pol in (0, 1):
                        ctx.align_flat(opt, pol, reads, refs, sc)
                ctx.score_ptrs(0, reads, refs, sc)
                ctx.align_ptrs(1, 0, reads, refs, sc)
            pr, ro = synth.pack_batch(reads)
            pf, fo = synth.pack_batch(refs)
            ctx.score_packed(0, pr, ro, pf, fo)
            for opt in (0, 1):
                ctx.align_packed(opt, 1 - opt, pr, ro, pf, fo)
    print("sanitize deck done")


if __name__ == "__main__":
    main()
