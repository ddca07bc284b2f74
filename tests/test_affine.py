"""Affine-gap (Gotoh) variant, SURVEY.md 8(f) rank 4: opt values 2 / 3 with the gap-open score in the bits above the
algorithm nibble (include/versalign_cuda.h).  The reference has no such kernel ("parity unpinned"): the checker is
oracle/va_oracle_affine.c, anchored to the pinned linear-gap oracle by the property that gap_open == 0 reproduces it.

  CPU : the property on the oracle (scores, end cells, every byte of the alignments)
  GPU : the CUDA path against the affine oracle -- flat, packed (CIGAR) and plug-in boundaries -- and the same property
"""
import numpy as np
import pytest

from oracle import binding as ora
from tests.helpers import used_region_equal
from versalignlib_b200 import synth

SCORINGS = [(2, -1, -3, -3), (3, -2, -1, -4), (1, 0, -7, -1), (5, -4, -2, -2)]
SW_AFFINE, NW_AFFINE = 2, 3


def affine_opt(opt: int, gap_open: int) -> int:
    """VA_OPT_*_AFFINE | VA_OPT_GAP_OPEN(gap_open) (include/versalign_cuda.h)"""
    return (SW_AFFINE if opt == ora.SW else NW_AFFINE) | (((-gap_open) & 0xFFFF) << 8)


def _decks():
    out = [("uniform", *synth.uniform_batch(400, 100, 150, p_sub=0.1, q_indel=0.05, seed=3)),
           ("mixed", *synth.mixed_batch(300, 30, 130, p_sub=0.1, q_indel=0.04, seed=4)[:2]),
           ("edge", *synth.edge_deck(48, 64)),
           ("tiny", *synth.uniform_batch(65, 17, 9, p_sub=0.3, seed=7)),
           ("random", *synth.uniform_batch(200, 64, 96, independent=True, seed=5))]
    r, f = synth.uniform_batch(300, 60, 90, p_sub=0.1, q_indel=0.05, seed=6)
    out.append(("dirty", synth.sprinkle(1, r, 0.03), synth.sprinkle(2, f, 0.03)))
    return out


DECKS = _decks()


def test_oracle_gap_open_zero_is_the_linear_oracle():
    for name, r, f in DECKS:
        for sc in SCORINGS:
            for opt in (ora.SW, ora.NW):
                assert np.array_equal(ora.score_affine(opt, r, f, sc, 0), ora.score(opt, r, f, sc)), (name, sc, opt)
                a, b, st, en, _ = ora.align_affine(opt, r, f, sc, 0)
                oa, ob, ost, oen = ora.align(opt, ora.POLICY_DEFAULT_OCL, r, f, sc)
                assert np.array_equal(st, ost) and np.array_equal(en, oen), (name, sc, opt)
                assert used_region_equal(a, b, st, oa, ob, ost).size == 0, (name, sc, opt)


def test_oracle_affine_gaps_cost_what_they_should():
    """One long gap beats two short ones exactly when the open score says so (hand-made case)."""
    read = np.frombuffer(b"ACGTACGTAAGGCCTTACGT", dtype=np.uint8)[None, :].copy()
    ref = np.frombuffer(b"ACGTACGTAATTTTGGCCTTACGT", dtype=np.uint8)[None, :].copy()  # the read with TTTT inserted
    a, b, st, en, s = ora.align_affine(ora.SW, read, ref, (2, -3, -1, -1), -4)
    L = a.shape[1]
    assert bytes(a[0, st[0]:L - 1]) == b"ACGTACGTAA----GGCCTTACGT" and bytes(b[0, st[0]:L - 1]) == bytes(ref[0])
    assert s[0] == 20 * 2 - 4 - 4 * 1


@pytest.mark.gpu
def test_cuda_affine_against_the_affine_oracle():
    from versalignlib_b200 import capi
    with capi.CudaContext(devices=[0]) as ctx:
        for name, r, f in DECKS:
            for sc in SCORINGS:
                for gap_open in (0, -2, -11):
                    for opt in (ora.SW, ora.NW):
                        o = affine_opt(opt, gap_open)
                        assert np.array_equal(ctx.score_flat(o, r, f, sc), ora.score_affine(opt, r, f, sc, gap_open)), (name, sc, gap_open, opt)
                        a, b, st, en = ctx.align_flat(o, 0, r, f, sc)
                        oa, ob, ost, oen, _ = ora.align_affine(opt, r, f, sc, gap_open)
                        assert np.array_equal(st, ost) and np.array_equal(en, oen), (name, sc, gap_open, opt)
                        assert used_region_equal(a, b, st, oa, ob, ost).size == 0, (name, sc, gap_open, opt)
                        if gap_open == 0:  # the anchor: the linear-gap modes, which run on other kernels (packed)
                            la, lb, lst, len_ = ctx.align_flat(opt, 0, r, f, sc)
                            assert np.array_equal(st, lst) and np.array_equal(en, len_) and np.array_equal(a, la) and np.array_equal(b, lb)
        # other containers and sizes: CIGARs through the packed boundary, a multi-chunk batch, a long pair
        name, r, f = DECKS[1]
        pr, ro = synth.pack_batch(r)
        pf, fo = synth.pack_batch(f)
        rr = np.ascontiguousarray(r[:, :max(int(np.diff(ro).max()), 1)])
        ff = np.ascontiguousarray(f[:, :max(int(np.diff(fo).max()), 1)])
        for opt in (ora.SW, ora.NW):
            scores, coords, coff, cigar = ctx.align_packed(affine_opt(opt, -6), 0, pr, ro, pf, fo)
            oa, ob, ost, oen, osc = ora.align_affine(opt, rr, ff, (2, -1, -3, -3), -6)
            wc, woff, wcig = synth.cigar_from_strings(oa, ob, ost, oen)
            assert np.array_equal(coords, wc) and np.array_equal(coff, woff) and np.array_equal(cigar, wcig), opt
            if opt == ora.SW:
                assert np.array_equal(scores, osc)
        r, f = synth.uniform_batch(90_000, 40, 52, p_sub=0.1, q_indel=0.05, seed=11)
        assert np.array_equal(ctx.score_flat(affine_opt(ora.SW, -4), r, f), ora.score_affine(ora.SW, r, f, (2, -1, -3, -3), -4))
        r, f = synth.uniform_batch(3, 900, 1300, p_sub=0.1, q_indel=0.04, seed=12)
        a, b, st, en = ctx.align_flat(affine_opt(ora.NW, -7), 0, r, f)
        oa, ob, ost, oen, _ = ora.align_affine(ora.NW, r, f, (2, -1, -3, -3), -7)
        assert np.array_equal(st, ost) and np.array_equal(en, oen) and used_region_equal(a, b, st, oa, ob, ost).size == 0


@pytest.mark.gpu
def test_affine_through_the_plugin_boundary():
    """opt 2 / 3 through dlopen + the virtual calls; the gap-open score is the optional key score_gap_open."""
    from versalignlib_b200 import capi
    from versalignlib_b200.host import PluginHost
    _, r, f = DECKS[0]
    sc = (2, -1, -3, -3)
    with PluginHost(capi.library_path(), r.shape[1], f.shape[1], sc, num_threads=4, extra={"cuda_devices": 1, "score_gap_open": -5}, verbosity=0) as h:
        h.stage(r, f, scattered=True)
        for opt, alg in ((ora.SW, SW_AFFINE), (ora.NW, NW_AFFINE)):
            assert np.array_equal(h.score_staged(alg), ora.score_affine(opt, r, f, sc, -5)), opt
            a, b, fields = h.align_staged(alg)
            oa, ob, ost, _, _ = ora.align_affine(opt, r, f, sc, -5)
            assert np.array_equal(fields[:, 0], ost) and used_region_equal(a, b, fields[:, 0], oa, ob, ost).size == 0, opt
        # the reference's own modes are untouched by the extra key
        assert np.array_equal(h.score_staged(ora.SW), ora.score(ora.SW, r, f, sc))
