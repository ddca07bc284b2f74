"""Parity of the CUDA path against the CPU oracle (which tests/test_oracle_vs_reference.py pins
against the reference's own kernels).  Every call goes through the C ABI or the plug-in boundary.
Bit-exact: scores, start offsets, end cells and every byte of the gapped strings."""
import numpy as np
import pytest

from oracle import binding as ora
from versalignlib_b200 import capi, synth
from versalignlib_b200.host import PluginHost
from tests.helpers import used_region_equal

pytestmark = pytest.mark.gpu

PARAM_SETS = [(2, -1, -3, -3), (3, -2, -1, -4), (5, -4, -1, -7), (1, 0, -7, -1), (2, -1, 0, -2)]


@pytest.fixture(scope="module")
def ctx():
    c = capi.CudaContext(devices=[0])
    yield c
    c.close()


def _batches():
    out = []
    r, f = synth.uniform_batch(1000, 100, 150, p_sub=0.10, seed=synth.BASE_SEED + 1)
    out.append(("c1_100x150", r, f))
    r, f = synth.uniform_batch(777, 150, 150, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 2)
    out.append(("c2_150x150", r, f))
    r, f = synth.uniform_batch(500, 64, 96, independent=True, seed=synth.BASE_SEED + 5)
    out.append(("c5_random64x96", r, f))
    r, f, _, _ = synth.mixed_batch(900, 30, 130, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 3)
    out.append(("mixed30-130", r, f))
    out.append(("dirty", synth.sprinkle(11, r, 0.03), synth.sprinkle(12, f, 0.03)))
    er, ef = synth.edge_deck(48, 64)
    out.append(("edge", er, ef))
    r, f = synth.uniform_batch(65, 17, 9, p_sub=0.3, seed=7)
    out.append(("tiny17x9", r, f))
    r, f = synth.uniform_batch(40, 250, 250, p_sub=0.1, q_indel=0.03, seed=8)
    out.append(("250x250", r, f))
    return out


BATCHES = _batches()


@pytest.mark.parametrize("opt", [ora.SW, ora.NW])
def test_scores_flat_and_ptrs(ctx, opt):
    for label, reads, refs in BATCHES:
        for sc in PARAM_SETS:
            want = ora.score(opt, reads, refs, sc)
            got = ctx.score_flat(opt, reads, refs, sc)
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, f"flat {label} opt={opt} sc={sc}: {bad[:5]} got={got[bad[:5]]} want={want[bad[:5]]}"
        got = ctx.score_ptrs(opt, reads, refs, PARAM_SETS[0])
        assert np.array_equal(got, ora.score(opt, reads, refs, PARAM_SETS[0])), f"ptrs {label}"


@pytest.mark.parametrize("policy", [ora.POLICY_DEFAULT_OCL, ora.POLICY_SIMD])
@pytest.mark.parametrize("opt", [ora.SW, ora.NW])
def test_alignments_flat(ctx, opt, policy):
    for label, reads, refs in BATCHES:
        for sc in PARAM_SETS[:4]:
            oa, ob, ostart, oend = ora.align(opt, policy, reads, refs, sc)
            a, b, start, end = ctx.align_flat(opt, policy, reads, refs, sc)
            assert np.array_equal(end, oend), f"{label} opt={opt} pol={policy} sc={sc}: end cells differ at {np.nonzero((end != oend).any(axis=1))[0][:5]}"
            bad = used_region_equal(a, b, start, oa, ob, ostart)
            assert bad.size == 0, f"{label} opt={opt} pol={policy} sc={sc}: pairs {bad[:5]}"
            # the flat entry point promises zeros before start and a NUL at L-1
            assert np.array_equal(a, oa) and np.array_equal(b, ob)


def test_alignments_ptrs(ctx):
    label, reads, refs = BATCHES[1]
    oa, ob, ostart, oend = ora.align(ora.NW, 0, reads, refs)
    a, b, start, end = ctx.align_ptrs(ora.NW, 0, reads, refs)
    assert np.array_equal(end, oend)
    assert used_region_equal(a, b, start, oa, ob, ostart).size == 0
    assert np.all(a[:, -1] == 0) and np.all(b[:, -1] == 0)


@pytest.mark.parametrize("policy", [0, 1])
def test_through_plugin_boundary(policy):
    """dlopen + set_parameters + set_logger + spawn + the two virtual calls, like the reference driver."""
    for label, reads, refs in BATCHES[:4]:
        sc = PARAM_SETS[1]
        with PluginHost(capi.library_path(), reads.shape[1], refs.shape[1], sc, num_threads=4,
                        extra={"cuda_traceback_policy": policy, "cuda_devices": 1}, verbosity=0) as h:
            h.stage(reads, refs, scattered=True)
            for opt in (ora.SW, ora.NW):
                got = h.score_staged(opt)
                assert np.array_equal(got, ora.score(opt, reads, refs, sc)), f"{label} opt={opt}"
                a, b, f = h.align_staged(opt)
                assert h.alignments_terminated()
                oa, ob, ostart, _ = ora.align(opt, policy, reads, refs, sc)
                L = reads.shape[1] + refs.shape[1]
                assert np.all(f[:, 0] == f[:, 2]) and np.all(f[:, 1] == L - 1) and np.all(f[:, 3] == L - 1)
                assert used_region_equal(a, b, f[:, 0], oa, ob, ostart).size == 0, f"{label} opt={opt}"


def test_unsupported_opt_touches_nothing(ctx):
    _, reads, refs = BATCHES[0]
    out = np.full(reads.shape[0], 1234, dtype=np.int16)
    ctx.score_flat(7, reads, refs, out=out)
    assert np.all(out == 1234)
    with PluginHost(capi.library_path(), reads.shape[1], refs.shape[1], extra={"cuda_devices": 1}, verbosity=0) as h:
        h.stage(reads, refs)
        sc = np.full(reads.shape[0], 77, dtype=np.int16)
        h.score_staged(5, out=sc)
        assert np.all(sc == 77)


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 129])
def test_ragged_batch_sizes(ctx, n):
    reads, refs = synth.uniform_batch(max(n, 1), 33, 47, p_sub=0.2, seed=100 + n)
    reads, refs = reads[:n], refs[:n]
    for opt in (ora.SW, ora.NW):
        assert np.array_equal(ctx.score_flat(opt, reads, refs), ora.score(opt, reads, refs) if n else np.zeros(0, np.int16))
        a, b, start, end = ctx.align_flat(opt, 0, reads, refs)
        if n:
            oa, ob, ostart, oend = ora.align(opt, 0, reads, refs)
            assert np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(start, ostart)


def test_device_resident_entry_points(ctx):
    import torch
    _, reads, refs = BATCHES[1]
    dev = torch.device("cuda:0")
    dr, df = torch.from_numpy(reads).to(dev), torch.from_numpy(refs).to(dev)
    n, L = reads.shape[0], reads.shape[1] + refs.shape[1]
    stream = torch.cuda.current_stream().cuda_stream
    for opt in (ora.SW, ora.NW):
        ds = torch.zeros(n, dtype=torch.int16, device=dev)
        ctx.score_device(opt, dr, df, ds, stream=stream)
        torch.cuda.synchronize()
        assert np.array_equal(ds.cpu().numpy(), ora.score(opt, reads, refs))
        da = torch.zeros((n, L), dtype=torch.uint8, device=dev)
        db = torch.zeros((n, L), dtype=torch.uint8, device=dev)
        dst = torch.zeros(n, dtype=torch.int16, device=dev)
        de = torch.zeros((n, 2), dtype=torch.int16, device=dev)
        ctx.align_device(opt, 0, dr, df, da, db, dst, de, stream=stream)
        torch.cuda.synchronize()
        oa, ob, ostart, oend = ora.align(opt, 0, reads, refs)
        assert np.array_equal(dst.cpu().numpy(), ostart) and np.array_equal(de.cpu().numpy(), oend)
        assert np.array_equal(da.cpu().numpy(), oa) and np.array_equal(db.cpu().numpy(), ob)


def test_large_batch_properties(ctx):
    """At a size the oracle cannot check pair by pair in seconds: size-independent properties.
    (1) identical pairs score match*len in both modes; (2) SW score is invariant under the
    '\\0' padding of the batch; (3) a sub-sample agrees with the oracle."""
    n = 200_000
    reads, refs = synth.uniform_batch(n, 100, 150, p_sub=0.1, seed=synth.BASE_SEED + 1)
    got = ctx.score_flat(ora.SW, reads, refs)
    idx = np.random.default_rng(0).choice(n, 3000, replace=False)
    assert np.array_equal(got[idx], ora.score(ora.SW, np.ascontiguousarray(reads[idx]), np.ascontiguousarray(refs[idx])))
    padded_r = np.zeros((n, 120), np.uint8); padded_r[:, :100] = reads
    padded_f = np.zeros((n, 160), np.uint8); padded_f[:, :150] = refs
    assert np.array_equal(ctx.score_flat(ora.SW, padded_r, padded_f), got)
    same = ctx.score_flat(ora.NW, refs, refs, (3, -2, -5, -5))
    assert np.all(same == 450)


def test_general_kernel_only_subprocess():
    """Same parity deck with the packed kernels switched off (VERSALIGN_CUDA_GENERAL_ONLY=1), in a
    fresh process because the switch is read once per process."""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np
from oracle import binding as ora
from versalignlib_b200 import capi, synth
reads, refs = synth.uniform_batch(600, 100, 150, p_sub=0.1, q_indel=0.02, seed=3)
with capi.CudaContext(devices=[0]) as ctx:
    for opt in (0, 1):
        assert np.array_equal(ctx.score_flat(opt, reads, refs), ora.score(opt, reads, refs))
        for pol in (0, 1):
            a, b, s, e = ctx.align_flat(opt, pol, reads, refs)
            oa, ob, os_, oe = ora.align(opt, pol, reads, refs)
            assert np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(s, os_) and np.array_equal(e, oe)
print("general-only ok")
'''
    env = dict(os.environ, VERSALIGN_CUDA_GENERAL_ONLY="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "general-only ok" in r.stdout, r.stdout + r.stderr


def test_wide_scores_fall_back_to_general_kernel(ctx):
    """Scores that do not fit the packed kernels' 8-bit tables / 16-bit headroom stay exact."""
    _, reads, refs = BATCHES[0]
    for sc in [(200, -150, -300, -300), (20, -30, -50, -40)]:
        for opt in (ora.SW, ora.NW):
            assert np.array_equal(ctx.score_flat(opt, reads, refs, sc), ora.score(opt, reads, refs, sc)), (sc, opt)
        a, b, start, end = ctx.align_flat(ora.NW, 0, reads, refs, sc)
        oa, ob, ostart, oend = ora.align(ora.NW, 0, reads, refs, sc)
        assert np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(start, ostart)


def test_c5_adversarial_parameter_grid(ctx):
    """BASELINE config 5: high-mismatch random pairs over the whole scoring grid, asymmetric gaps
    included -- scores and alignments (both pointer policies) bit-exact against the oracle."""
    decks = [synth.uniform_batch(192, 64, 96, independent=True, seed=synth.BASE_SEED + 5),
             synth.uniform_batch(128, 150, 150, p_sub=0.5, seed=synth.BASE_SEED + 55)]
    combos = [(m, x, gr, gf) for m in (1, 2, 5) for x in (0, -1, -4) for gr in (-1, -3, -7) for gf in (-1, -3, -7)]
    assert len(combos) == 81
    for reads, refs in decks:
        for ci, sc in enumerate(combos):
            for opt in (ora.SW, ora.NW):
                assert np.array_equal(ctx.score_flat(opt, reads, refs, sc), ora.score(opt, reads, refs, sc)), (sc, opt)
                pol = ci & 1  # alternate policies over the grid; both are covered for every gap pair
                a, b, start, end = ctx.align_flat(opt, pol, reads, refs, sc)
                oa, ob, ostart, oend = ora.align(opt, pol, reads, refs, sc)
                assert np.array_equal(start, ostart) and np.array_equal(end, oend), (sc, opt, pol)
                assert np.array_equal(a, oa) and np.array_equal(b, ob), (sc, opt, pol)


def test_long_pairs_intra_task_kernel(ctx):
    """Few, long pairs take the warp-per-pair wavefront kernel (va_intra.cu); the oracle is the only
    reference for it (the reference library has no intra-task kernel)."""
    for n, rl, fl, seed in [(40, 1000, 1200, 31), (6, 3000, 3500, 32), (9, 700, 2049, 33)]:
        reads, refs = synth.uniform_batch(n, rl, fl, p_sub=0.10, q_indel=0.03, seed=seed)
        for sc in [(2, -1, -3, -3), (3, -2, -1, -4)]:
            got = ctx.score_flat(ora.SW, reads, refs, sc)
            want = ora.score(ora.SW, reads, refs, sc)
            assert np.array_equal(got, want), (n, rl, fl, sc, np.nonzero(got != want)[0][:5])
    # mixed lengths + an odd pair count: partial last pass, general-kernel tail
    reads, refs, _, _ = synth.mixed_batch(33, 900, 1500, p_sub=0.1, q_indel=0.02, seed=34)
    assert np.array_equal(ctx.score_flat(ora.SW, reads, refs), ora.score(ora.SW, reads, refs))


@pytest.mark.parametrize("opt", [ora.SW, ora.NW])
def test_packed_entry_points(ctx, opt):
    """Batch-friendly containers, same kernels: offset-addressed sequences in; scores, sequence coordinates
    and CIGARs out must be exactly what the reference-style outputs (oracle) say."""
    for name, reads, refs in [BATCHES[3], BATCHES[1], BATCHES[6], BATCHES[2]]:  # mixed lengths, uniform with indels, tiny, random (raw-move fallback)
        pr, ro = synth.pack_batch(reads)
        pf, fo = synth.pack_batch(refs)
        # the semantics are the reference's on the batch padded to ITS maximum lengths
        reads = np.ascontiguousarray(reads[:, :max(int(np.diff(ro).max()), 1)])
        refs = np.ascontiguousarray(refs[:, :max(int(np.diff(fo).max()), 1)])
        assert np.array_equal(ctx.score_packed(opt, pr, ro, pf, fo), ora.score(opt, reads, refs)), name
        for policy in (0, 1):
            scores, coords, coff, cigar = ctx.align_packed(opt, policy, pr, ro, pf, fo)
            oa, ob, ostart, oend = ora.align(opt, policy, reads, refs)
            want_coords, want_off, want_cigar = synth.cigar_from_strings(oa, ob, ostart, oend)
            assert np.array_equal(coff, want_off), (name, policy)
            assert np.array_equal(cigar, want_cigar), (name, policy)
            assert np.array_equal(coords, want_coords), (name, policy)
            if opt == ora.SW:
                assert np.array_equal(scores, ora.score(opt, reads, refs)), (name, policy)


def test_packed_entry_points_gap_rich_alignments(ctx):
    """Cheap gaps and expensive mismatches give alignments with far more CIGAR runs than move slots: every size
    class of the device-side compaction -- fits the estimate, needs the second copy, exceeds the pinned block,
    exceeds the device block (moves replayed on the host) -- must return the oracle's CIGARs (regression: the pinned
    block is rounded differently from the device block and the zone between the two failed with 'invalid argument')."""
    for n, rl, fl, seed in [(3000, 64, 96, 21), (20_000, 40, 40, 22), (70_000, 24, 31, 23)]:
        reads, refs = synth.uniform_batch(n, rl, fl, independent=True, seed=seed)
        pr, ro = synth.pack_batch(reads)
        pf, fo = synth.pack_batch(refs)
        for sc in [(2, -30, -1, -1), (2, -1, -3, -3), (5, -4, -1, -7)]:
            for opt in (ora.SW, ora.NW):
                scores, coords, coff, cigar = ctx.align_packed(opt, 0, pr, ro, pf, fo, sc)
                want_coords, want_off, want_cigar = synth.cigar_from_strings(*ora.align(opt, 0, reads, refs, sc))
                assert np.array_equal(coff, want_off) and np.array_equal(cigar, want_cigar) and np.array_equal(coords, want_coords), (n, sc, opt)


def test_positive_gap_scores_stay_exact(ctx):
    """Gap scores > 0 are outside the packed kernels' domain (their padding arguments need gaps <= 0):
    the general kernel must take over, results unchanged."""
    _, reads, refs = BATCHES[3]
    for sc in [(2, -1, 1, -3), (2, -1, -2, 1)]:
        for opt in (ora.SW, ora.NW):
            assert np.array_equal(ctx.score_flat(opt, reads, refs, sc), ora.score(opt, reads, refs, sc)), (sc, opt)


def test_mixed_lengths_at_scale(ctx):
    """Length-bucketed mixed batch big enough that every strip width, padded last strips, and the
    solo path (odd leftovers of the bucketing) all occur many times: all four functions against the oracle."""
    reads, refs, _, _ = synth.mixed_batch(12_000, 60, 250, p_sub=0.1, q_indel=0.02, seed=synth.BASE_SEED + 33)
    for opt in (ora.SW, ora.NW):
        assert np.array_equal(ctx.score_flat(opt, reads, refs), ora.score(opt, reads, refs)), opt
        a, b, start, end = ctx.align_flat(opt, 0, reads, refs)
        oa, ob, ostart, oend = ora.align(opt, 0, reads, refs)
        assert np.array_equal(start, ostart) and np.array_equal(end, oend), opt
        assert used_region_equal(a, b, start, oa, ob, ostart).size == 0, opt


def test_workspace_reuse_across_scorings(ctx):
    """The device workspace is reused between calls.  A call whose values are large (NW with big scores)
    must not leak into a later SW-align call through the halves of the words that solo threads share
    (regression: SW align's packed key crosses lanes)."""
    r0, f0 = synth.uniform_batch(1000, 100, 150, p_sub=0.10, seed=synth.BASE_SEED + 1)
    big = (20, -30, -50, -40)
    assert np.array_equal(ctx.score_flat(ora.NW, r0, f0, big), ora.score(ora.NW, r0, f0, big))
    ctx.align_flat(ora.NW, 0, r0, f0, big)
    _, reads, refs = BATCHES[3]
    for opt in (ora.SW, ora.NW):
        a, b, start, end = ctx.align_flat(opt, 0, reads, refs)
        oa, ob, ostart, oend = ora.align(opt, 0, reads, refs)
        assert np.array_equal(start, ostart) and np.array_equal(end, oend), opt
        assert used_region_equal(a, b, start, oa, ob, ostart).size == 0, opt


def test_fasta_to_packed_kernels(ctx, tmp_path):
    """The path a maintainer would wire for a file-driven run: FASTA -> va_fasta_load -> packed entry points,
    against the oracle on the same sequences padded the reference's way (parse_fasta + pad())."""
    _, reads, refs = BATCHES[3]
    paths = []
    for tag, arr in (("reads", reads), ("refs", refs)):
        path = str(tmp_path / f"{tag}.fa")
        with open(path, "wb") as f:
            for i in range(arr.shape[0]):
                seq = arr[i].tobytes().rstrip(b"\0")
                f.write(b">%s%d\n" % (tag.encode(), i))
                for o in range(0, len(seq), 70):
                    f.write(seq[o:o + 70] + b"\n")
        paths.append(path)
    pr, ro, rmax = capi.fasta_load(paths[0])
    pf, fo, fmax = capi.fasta_load(paths[1])
    padded_r = np.ascontiguousarray(reads[:, :rmax])
    padded_f = np.ascontiguousarray(refs[:, :fmax])
    for opt in (ora.SW, ora.NW):
        assert np.array_equal(ctx.score_packed(opt, pr, ro, pf, fo), ora.score(opt, padded_r, padded_f)), opt
        scores, coords, coff, cigar = ctx.align_packed(opt, 0, pr, ro, pf, fo)
        want = synth.cigar_from_strings(*ora.align(opt, 0, padded_r, padded_f))
        assert np.array_equal(coords, want[0]) and np.array_equal(coff, want[1]) and np.array_equal(cigar, want[2]), opt


def test_long_pairs_warp_per_pair_general_kernel(ctx):
    """Long pairs outside the packed kernels' 16-bit domain (alignments with large scores, the SSE/AVX policy,
    wide scoring, N inside the ref) take the warp-per-pair general kernel (fill_general_intra_kernel): all
    four functions, both policies, odd sizes, several 512-column passes and a partial last one."""
    reads, refs = synth.uniform_batch(9, 1100, 1300, p_sub=0.10, q_indel=0.03, seed=41)
    dirty_refs = synth.sprinkle(5, refs, 0.01)
    mixed_r, mixed_f, _, _ = synth.mixed_batch(7, 700, 1400, p_sub=0.1, q_indel=0.02, seed=42)
    for name, r, f in (("uniform", reads, refs), ("dirty", reads, dirty_refs), ("mixed", mixed_r, mixed_f)):
        for sc in [(2, -1, -3, -3), (20, -30, -50, -20)]:  # 20 x 1100 stays inside int16 (the exactness domain)
            for opt in (ora.SW, ora.NW):
                assert np.array_equal(ctx.score_flat(opt, r, f, sc), ora.score(opt, r, f, sc)), (name, sc, opt)
                for policy in (0, 1):
                    a, b, start, end = ctx.align_flat(opt, policy, r, f, sc)
                    oa, ob, ostart, oend = ora.align(opt, policy, r, f, sc)
                    assert np.array_equal(start, ostart) and np.array_equal(end, oend), (name, sc, opt, policy)
                    assert used_region_equal(a, b, start, oa, ob, ostart).size == 0, (name, sc, opt, policy)


def test_long_pairs_intra_task_align(ctx):
    """Few, long pairs: the packed intra-task kernels (a CTA per pair-of-pairs, va_intra.cu) fill every mode and
    the warp-per-pair traceback walks their [duo][strip][row] direction words.  All four functions, both NW pointer
    policies, equal and unequal gap scores, several 512-column passes with a partial last one, duos whose reads and
    refs differ in length, a dirty ref (that pair falls to the general kernel inside the same chunk), an odd pair
    count -- through the flat (fixed-stride strings), scattered (compact strings) and packed (CIGAR) entry points."""
    decks = []
    r, f = synth.uniform_batch(12, 1000, 1200, p_sub=0.10, q_indel=0.03, seed=51)
    decks.append(("1000x1200", r, f))
    r, f = synth.uniform_batch(5, 2600, 3100, p_sub=0.12, q_indel=0.03, seed=52)
    decks.append(("2600x3100_odd", r, f))
    r, f, _, _ = synth.mixed_batch(14, 600, 1500, p_sub=0.1, q_indel=0.02, seed=53)
    decks.append(("mixed600-1500", r, f))
    r, f = synth.uniform_batch(8, 300, 1024, independent=True, seed=54)  # unrelated sequences: short local alignments
    decks.append(("random300x1024", r, f))
    r, f, rl_, fl_ = synth.mixed_batch(18, 1, 1200, p_sub=0.1, q_indel=0.02, seed=56)  # reads / refs of 1, 2, 3 ... bases next to long ones
    r[0, 1:] = 0; r[1, 2:] = 0; r[2, 3:] = 0; f[3, 1:] = 0; r[3, 1:] = 0; f[4, 17:] = 0; r[4, 16:] = 0
    decks.append(("mixed1-1200", r, f))
    r, f = synth.uniform_batch(6, 1100, 1300, p_sub=0.10, q_indel=0.03, seed=55)
    f = f.copy()
    f[2, 400] = ord("N")
    decks.append(("one_dirty_ref", r, f))
    for name, r, f in decks:
        for sc in [(2, -1, -3, -3), (3, -2, -1, -4), (1, -1, -2, -2)]:
            for opt in (ora.SW, ora.NW):
                assert np.array_equal(ctx.score_flat(opt, r, f, sc), ora.score(opt, r, f, sc)), (name, sc, opt)
                for policy in (0, 1):
                    oa, ob, ostart, oend = ora.align(opt, policy, r, f, sc)
                    a, b, start, end = ctx.align_flat(opt, policy, r, f, sc)
                    assert np.array_equal(start, ostart) and np.array_equal(end, oend), (name, sc, opt, policy)
                    assert used_region_equal(a, b, start, oa, ob, ostart).size == 0, (name, sc, opt, policy)
    # the other two result containers on one deck
    name, r, f = decks[0]
    for opt in (ora.SW, ora.NW):
        oa, ob, ostart, oend = ora.align(opt, 0, r, f)
        a, b, start, end = ctx.align_ptrs(opt, 0, r, f)
        assert np.array_equal(start, ostart) and used_region_equal(a, b, start, oa, ob, ostart).size == 0, opt
        pr, ro = synth.pack_batch(r)
        pf, fo = synth.pack_batch(f)
        scores, coords, coff, cigar = ctx.align_packed(opt, 0, pr, ro, pf, fo)
        wc, woff, wcig = synth.cigar_from_strings(oa, ob, ostart, oend)
        assert np.array_equal(coords, wc) and np.array_equal(coff, woff) and np.array_equal(cigar, wcig), opt


def test_nw_align_end_aligned_duos(ctx):
    """Packed NW align takes duos whose reads differ in length by starting the shorter lane late (CODE_PRE
    rows): refs of one length, reads of every length from 1 up, odd and even offsets, trimmed and full refs."""
    rng = np.random.default_rng(77)
    n, L = 600, 96
    refs = synth.random_seqs(rng, n, L)
    reads = synth.mutate_from_ref(rng, refs, L, 0.1, 0.03)
    lens = rng.integers(1, L + 1, size=n)
    reads = np.where(np.arange(L)[None, :] < lens[:, None], reads, 0).astype(np.uint8)
    short_refs = refs.copy()
    short_refs[:, 80:] = 0  # padded refs: the pad-column rule reads the last true column at shifted rows
    for f in (refs, short_refs):
        for sc in PARAM_SETS[:3]:
            for pol in (0, 1):  # both pointer policies run on the packed kernel (second plane: UP >= LEFT / LEFT >= UP)
                a, b, start, end = ctx.align_flat(ora.NW, pol, reads, f, sc)
                oa, ob, ostart, oend = ora.align(ora.NW, pol, reads, f, sc)
                assert np.array_equal(start, ostart) and np.array_equal(end, oend), (sc, pol)
                assert used_region_equal(a, b, start, oa, ob, ostart).size == 0, (sc, pol)


def test_c2_workload_sample_at_scale(ctx):
    """The bench workload (C2: NW compute_alignments, 150 x 150) at a size that runs several chunks through the
    host pipeline: a random sample of pairs must carry exactly the oracle's alignments, and the flat and the
    packed/CIGAR results must describe the same paths."""
    n = 300_000
    reads, refs = synth.uniform_batch(n, 150, 150, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 2)
    a, b, start, end = ctx.align_flat(ora.NW, 0, reads, refs)
    idx = np.sort(np.random.default_rng(1).choice(n, 4000, replace=False))
    oa, ob, ostart, oend = ora.align(ora.NW, 0, np.ascontiguousarray(reads[idx]), np.ascontiguousarray(refs[idx]))
    assert np.array_equal(start[idx], ostart) and np.array_equal(end[idx], oend)
    assert used_region_equal(a[idx], b[idx], start[idx], oa, ob, ostart).size == 0
    pr, ro = synth.pack_batch(reads)
    pf, fo = synth.pack_batch(refs)
    scores, coords, coff, cigar = ctx.align_packed(ora.NW, 0, pr, ro, pf, fo)
    L = a.shape[1]
    moves = np.add.reduceat((cigar >> 4).astype(np.int64), coff[:-1]) if cigar.size else np.zeros(n, np.int64)
    assert np.array_equal(moves, L - 1 - start.astype(np.int64))
    assert np.array_equal(coords[:, 1], end[:, 0].astype(np.int32) + 1) and np.array_equal(coords[:, 3], end[:, 1].astype(np.int32) + 1)


@pytest.mark.parametrize("rl,fl", [(700, 700), (735, 740), (745, 745), (760, 775)])
def test_traceback_queue_shared_to_global_boundary(ctx, rl, fl):
    """Around read+ref = 1.5 k the traceback's move queues stop fitting into shared memory next to the
    kernel's static arrays and move to global memory (regression: the switch ignored the static part)."""
    reads, refs = synth.uniform_batch(70, rl, fl, p_sub=0.1, q_indel=0.02, seed=rl)
    for opt in (ora.SW, ora.NW):
        a, b, start, end = ctx.align_flat(opt, 0, reads, refs)
        oa, ob, ostart, oend = ora.align(opt, 0, reads, refs)
        assert np.array_equal(start, ostart) and np.array_equal(end, oend), opt
        assert used_region_equal(a, b, start, oa, ob, ostart).size == 0, opt


def test_align_alloc_allocator_boundary(ctx):
    """va_cuda_align_alloc: result blocks come from the caller's allocator on the staging threads; same
    alignments as the oracle; an allocator that runs dry fails the call with VA_ERR_MEMORY."""
    _, reads, refs = BATCHES[1]
    for opt in (ora.SW, ora.NW):
        a, b, start, end = ctx.align_alloc(opt, 0, reads, refs)
        oa, ob, ostart, oend = ora.align(opt, 0, reads, refs)
        assert np.array_equal(start, ostart) and np.array_equal(end, oend), opt
        assert used_region_equal(a, b, start, oa, ob, ostart).size == 0, opt
    with pytest.raises(capi.CudaError, match="allocator"):
        ctx.align_alloc(ora.NW, 0, reads, refs, fail_after=100)
    # the context stays usable after the failed call
    assert np.array_equal(ctx.score_flat(ora.SW, reads, refs), ora.score(ora.SW, reads, refs))


def test_mismatch_score_above_match(ctx):
    """A 'mismatch' score above the match score makes cells grow by the mismatch score: the packed kernels'
    16-bit range checks must use the larger of the two (else the general kernel takes the call)."""
    r, f = synth.uniform_batch(300, 250, 250, independent=True, seed=91)
    for sc in [(1, 4, -3, -3), (2, 5, -1, -2), (1, 120, -3, -3)]:
        for opt in (ora.SW, ora.NW):
            assert np.array_equal(ctx.score_flat(opt, r, f, sc), ora.score(opt, r, f, sc)), (sc, opt)
            a, b, start, end = ctx.align_flat(opt, 0, r, f, sc)
            oa, ob, ostart, oend = ora.align(opt, 0, r, f, sc)
            assert np.array_equal(start, ostart) and np.array_equal(end, oend), (sc, opt)
            assert used_region_equal(a, b, start, oa, ob, ostart).size == 0, (sc, opt)


def test_packed_inputs_in_pinned_memory(ctx):
    """Offset-addressed inputs in page-locked memory are read in place by the copy engines (no staging copy):
    same results as from pageable memory, on a batch big enough for several chunks and both devices' paths."""
    import torch
    reads, refs, _, _ = synth.mixed_batch(150_000, 60, 150, p_sub=0.08, q_indel=0.02, seed=77)
    pr, ro = synth.pack_batch(reads)
    pf, fo = synth.pack_batch(refs)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    want = ctx.align_packed(ora.NW, 0, pr, ro, pf, fo)
    got = ctx.align_packed(ora.NW, 0, pin(pr), pin(ro), pin(pf), pin(fo))
    for w, g in zip(want, got):
        assert np.array_equal(w, g)
    assert np.array_equal(ctx.score_packed(ora.SW, pin(pr), pin(ro), pin(pf), pin(fo)), ctx.score_packed(ora.SW, pr, ro, pf, fo))
    # against the oracle on a sample
    idx = np.arange(0, reads.shape[0], 37)
    oa, ob, ostart, oend = ora.align(ora.NW, 0, np.ascontiguousarray(reads[idx]), np.ascontiguousarray(refs[idx]))
    wc, woff, wcig = synth.cigar_from_strings(oa, ob, ostart, oend)
    scores, coords, coff, cigar = got
    assert np.array_equal(coords[idx], wc)
    for k, i in enumerate(idx[:500]):
        assert np.array_equal(cigar[coff[i]:coff[i + 1]], wcig[woff[k]:woff[k + 1]]), i


def test_c2_full_million_pairs_against_oracle(ctx):
    """The whole C2 bench workload, not a sample: 1 M x (150 vs 150) NW compute_alignments, every start index, end
    cell and byte of both gapped strings against the OpenMP oracle (about 15 s of host time on the GPU box)."""
    n = 1_000_000
    reads, refs = synth.uniform_batch(n, 150, 150, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 2)
    a, b, start, end = ctx.align_flat(ora.NW, 0, reads, refs)
    bad = 0
    for lo in range(0, n, 250_000):  # slices bound the oracle's output arrays
        hi = lo + 250_000
        oa, ob, ostart, oend = ora.align(ora.NW, 0, reads[lo:hi], refs[lo:hi])
        assert np.array_equal(start[lo:hi], ostart) and np.array_equal(end[lo:hi], oend), lo
        bad += used_region_equal(a[lo:hi], b[lo:hi], start[lo:hi], oa, ob, ostart).size
    assert bad == 0


def test_in_process_multi_device_shard():
    """One context over every visible device: the pair range is sharded inside the C ABI (one host thread per device,
    no exchange), results against the oracle -- flat, scattered and packed entry points.  Needs >= 2 GPUs."""
    nd = capi.device_count()
    if nd < 2:
        pytest.skip("needs at least 2 CUDA devices")
    reads, refs, rl, fl = synth.mixed_batch(150_001, 20, 120, p_sub=0.1, q_indel=0.02, seed=9)
    with capi.CudaContext(devices=list(range(nd))) as mctx:
        for opt in (ora.SW, ora.NW):
            assert np.array_equal(mctx.score_flat(opt, reads, refs), ora.score(opt, reads, refs)), ("score", opt)
            a, b, start, end = mctx.align_flat(opt, 0, reads, refs)
            oa, ob, ostart, oend = ora.align(opt, 0, reads, refs)
            assert np.array_equal(start, ostart) and np.array_equal(end, oend), ("align", opt)
            assert used_region_equal(a, b, start, oa, ob, ostart).size == 0, ("strings", opt)
        assert mctx.timings()["devices"] == nd
        pr, ro = synth.pack_batch(reads, rl)
        pf, fo = synth.pack_batch(refs, fl)
        tr, tf = np.ascontiguousarray(reads[:, :int(rl.max())]), np.ascontiguousarray(refs[:, :int(fl.max())])
        assert np.array_equal(mctx.score_packed(ora.SW, pr, ro, pf, fo), ora.score(ora.SW, tr, tf))
        scores, coords, coff, cigar = mctx.align_packed(ora.NW, 0, pr, ro, pf, fo)
        idx = np.arange(0, reads.shape[0], 37)
        want = synth.cigar_from_strings(*ora.align(ora.NW, 0, np.ascontiguousarray(tr[idx]), np.ascontiguousarray(tf[idx])))
        assert np.array_equal(coords[idx], want[0])
        got_runs = [cigar[coff[i]:coff[i + 1]].tolist() for i in idx]
        want_runs = [want[2][want[1][k]:want[1][k + 1]].tolist() for k in range(len(idx))]
        assert got_runs == want_runs


def test_device_resident_large_call_into_dirty_blocks(ctx):
    """A device-resident align call of more than one wave of the packed kernels (whole wave + 20 k pairs), both
    algorithms and pointer policies, garbage in the output blocks beforehand: the bytes before start[i] must come back
    zero (the blocks are cleared on the side stream while the fill kernels run, va_cabi.cu) and every byte must match
    the oracle, with and without the per-phase profiling events."""
    import torch
    dev = torch.device("cuda:0")
    wave = torch.cuda.get_device_properties(0).multi_processor_count * 1024
    n = wave + 20_000
    reads, refs = synth.uniform_batch(n, 100, 120, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 41)
    dr, df = torch.from_numpy(reads).to(dev), torch.from_numpy(refs).to(dev)
    L = reads.shape[1] + refs.shape[1]
    stream = torch.cuda.current_stream().cuda_stream
    for opt, policy in ((ora.NW, 0), (ora.SW, 0), (ora.NW, 1)):
        oa, ob, ostart, oend = ora.align(opt, policy, reads, refs)
        for profiled in (False, True):
            da = torch.full((n, L), 0x55, dtype=torch.uint8, device=dev)
            db = torch.full((n, L), 0x55, dtype=torch.uint8, device=dev)
            dst = torch.zeros(n, dtype=torch.int16, device=dev)
            de = torch.zeros((n, 2), dtype=torch.int16, device=dev)
            ctx.set_profiling(profiled)
            ctx.align_device(opt, policy, dr, df, da, db, dst, de, stream=stream)
            torch.cuda.synchronize()
            ctx.set_profiling(False)
            assert np.array_equal(dst.cpu().numpy(), ostart) and np.array_equal(de.cpu().numpy(), oend), (opt, policy, profiled)
            assert np.array_equal(da.cpu().numpy(), oa) and np.array_equal(db.cpu().numpy(), ob), (opt, policy, profiled)


def test_nw_align_plane_form_in_a_fresh_process():
    """The packed NW align kernel has two forms (va_nw.cu): tagged lanes (4V + tag, the default whenever 4x the value range
    fits 16 bits) and two bit planes.  The other parity tests run the tagged form; here the plane form takes the same
    kinds of deck (VERSALIGN_CUDA_NO_INBAND=1, read once per process): uniform, mixed-length (end-aligned duos, padded
    last strips, solo slots) and dirty, both pointer policies, flat and packed entry points."""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np
from oracle import binding as ora
from versalignlib_b200 import capi, synth
decks = [synth.uniform_batch(3000, 150, 150, p_sub=0.08, q_indel=0.02, seed=5)]
r, f, _, _ = synth.mixed_batch(4000, 30, 170, p_sub=0.08, q_indel=0.02, seed=6)
decks.append((r, f))
decks.append((synth.sprinkle(13, r, 0.02), synth.sprinkle(14, f, 0.02)))
with capi.CudaContext(devices=[0]) as ctx:
    for reads, refs in decks:
        for pol in (0, 1):
            a, b, s, e = ctx.align_flat(1, pol, reads, refs)
            oa, ob, os_, oe = ora.align(1, pol, reads, refs)
            assert np.array_equal(s, os_) and np.array_equal(e, oe)
            assert np.array_equal(a, oa) and np.array_equal(b, ob)
print("plane form ok")
'''
    env = dict(os.environ, VERSALIGN_CUDA_NO_INBAND="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "plane form ok" in r.stdout, r.stdout + r.stderr


def test_nw_align_scores_outside_the_tagged_range(ctx):
    """Scorings whose 4x value range or 4s' + 2 table entries leave the tagged form's 16-bit lanes / 8-bit tables run the
    plane form of the packed kernel in the same process, next to calls that run the tagged form."""
    reads, refs = synth.uniform_batch(2000, 150, 150, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 51)
    long_r, long_f = synth.uniform_batch(64, 800, 800, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 52)
    for (r, f), sc in (((reads, refs), (25, -4, -3, -5)),      # 4 * (25 + 8) + 2 does not fit a signed byte
                       ((reads, refs), (2, -1, -3, -3)),       # tagged
                       ((long_r, long_f), (6, -2, -3, -3)),    # 4 * (6 * 800 + 6 * 802) leaves 16 bits; V itself fits
                       ((long_r, long_f), (2, -1, -1, -1))):   # tagged again, longer strips
        for pol in (0, 1):
            a, b, start, end = ctx.align_flat(ora.NW, pol, r, f, sc)
            oa, ob, ostart, oend = ora.align(ora.NW, pol, r, f, sc)
            assert np.array_equal(start, ostart) and np.array_equal(end, oend), (sc, pol)
            assert np.array_equal(a, oa) and np.array_equal(b, ob), (sc, pol)


def test_scattered_pointers_with_long_sequences(ctx):
    """The gather of scattered sequences goes through a line buffer (va_cabi.cu, LineStreamer): sequences longer than the
    buffer take its direct path, short and long ones in one call order must not swap bytes.  9 k / 4.2 k / 130 base reads
    through the pointer entry points (what the plug-in calls), scores and alignments against the oracle."""
    for rl, fl, n in ((9000, 9500, 6), (4200, 300, 40), (130, 8300, 24), (8180, 4090, 20), (4097, 4095, 33)):
        reads, refs = synth.uniform_batch(n, rl, fl, p_sub=0.1, q_indel=0.02, seed=synth.BASE_SEED + 61 + rl)
        for opt in (ora.SW, ora.NW):
            assert np.array_equal(ctx.score_ptrs(opt, reads, refs), ora.score(opt, reads, refs)), (rl, fl, opt)
        a, b, start, end = ctx.align_ptrs(ora.NW, 0, reads, refs)
        oa, ob, ostart, oend = ora.align(ora.NW, 0, reads, refs)
        assert np.array_equal(start, ostart) and np.array_equal(end, oend), (rl, fl)
        assert used_region_equal(a, b, start, oa, ob, ostart).size == 0, (rl, fl)
