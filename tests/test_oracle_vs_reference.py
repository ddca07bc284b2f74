"""Pin the CPU oracle against the reference's OWN kernels, executed.

oracle/Makefile compiles libDefaultKernel.so / libSSEKernel.so / libAVXKernel.so straight
from /root/reference into oracle/_ref (git-ignored, travels to the GPU box).  They are
loaded through the reference's plug-in boundary by csrc/plugin_host.cpp.  Caveats of the
reference that shape these tests:
  * Default stores only the LOW BYTE of a score (memset(...,1), DefaultKernel.cpp:137,199);
  * SSE is racy with num_threads > 1 (SSEKernel.cpp:77-82) -> always 1 thread here;
  * SSE/AVX compute_alignments tails (n % 8, n % 16) return dangling pointers -> n % 16 == 0.
"""
import numpy as np
import pytest

from oracle import binding as ora
from versalignlib_b200 import synth
from versalignlib_b200.host import PluginHost
from tests.helpers import used_region_equal

pytestmark = pytest.mark.refso

PARAM_SETS = [(2, -1, -3, -3), (3, -2, -1, -4), (5, -4, -1, -7), (1, 0, -7, -1)]


def _for_default(reads, refs):
    """Default indexes char_to_score[] with a signed char (DefaultKernel.cpp:106): bytes
    >= 0x80 read out of bounds there (undefined behaviour), so that kernel never sees them.
    SSE/AVX mask with 0xDF and treat them as non-ACGT, which is what the oracle does."""
    return np.where(reads >= 0x80, ord("X"), reads).astype(np.uint8), np.where(refs >= 0x80, ord("X"), refs).astype(np.uint8)


def _need(name):
    p = ora.ref_lib(name)
    if p is None:
        pytest.skip(f"oracle/_ref/lib{name}Kernel.so not built (reference tree absent)")
    return p


def _batches():
    """(label, reads, refs): uniform, mixed-length padded, and dirty (N / lower case / junk)."""
    out = []
    r, f = synth.uniform_batch(320, 100, 150, p_sub=0.10, seed=synth.BASE_SEED + 1)
    out.append(("uniform100x150", r, f))
    r, f = synth.uniform_batch(320, 64, 96, independent=True, seed=synth.BASE_SEED + 5)
    out.append(("random64x96", r, f))
    r, f, _, _ = synth.mixed_batch(320, 40, 120, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 3)
    out.append(("mixed40-120", r, f))
    r2, f2 = synth.sprinkle(11, r, 0.02), synth.sprinkle(12, f, 0.02)
    out.append(("dirty", r2, f2))
    r, f = synth.uniform_batch(160, 150, 150, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 2)
    out.append(("c2_150x150", r, f))
    er, ef = synth.edge_deck(48, 64)
    reps = 32 // len(er) + 1
    er, ef = np.tile(er, (reps, 1))[:32], np.tile(ef, (reps, 1))[:32]
    out.append(("edge", np.ascontiguousarray(er), np.ascontiguousarray(ef)))
    return out


BATCHES = _batches()


@pytest.mark.parametrize("kernel", ["SSE", "AVX"])
@pytest.mark.parametrize("opt", [ora.SW, ora.NW])
def test_scores_match_simd_kernels(kernel, opt):
    lib = _need(kernel)
    for label, reads, refs in BATCHES:
        for sc in PARAM_SETS:
            with PluginHost(lib, reads.shape[1], refs.shape[1], sc, num_threads=1, verbosity=0) as h:
                got = h.score_alignments(opt, reads, refs)
            want = ora.score(opt, reads, refs, sc)
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, f"{kernel} {label} opt={opt} sc={sc}: {bad[:5]} ref={got[bad[:5]]} oracle={want[bad[:5]]}"


@pytest.mark.parametrize("opt", [ora.SW, ora.NW])
def test_scores_match_default_low_byte(opt):
    lib = _need("Default")
    for label, reads, refs in BATCHES:
        reads, refs = _for_default(reads, refs)
        for sc in PARAM_SETS:
            with PluginHost(lib, reads.shape[1], refs.shape[1], sc, num_threads=2, verbosity=0) as h:
                got = h.score_alignments(opt, reads, refs)  # array pre-zeroed; only byte 0 is written
            want = ora.score(opt, reads, refs, sc)
            assert np.array_equal(got.view(np.uint8)[0::2], want.view(np.uint8)[0::2]), f"{label} opt={opt} sc={sc}"


@pytest.mark.parametrize("kernel,policy", [("Default", ora.POLICY_DEFAULT_OCL), ("SSE", ora.POLICY_SIMD),
                                           ("AVX", ora.POLICY_SIMD)])
@pytest.mark.parametrize("opt", [ora.SW, ora.NW])
def test_alignments_match(kernel, policy, opt):
    lib = _need(kernel)
    for label, reads, refs in BATCHES:
        if kernel == "Default":
            reads, refs = _for_default(reads, refs)
        for sc in PARAM_SETS[:3]:
            with PluginHost(lib, reads.shape[1], refs.shape[1], sc, num_threads=1, verbosity=0) as h:
                a, b, f = h.compute_alignments(opt, reads, refs)
            oa, ob, ostart, _ = ora.align(opt, policy, reads, refs, sc)
            L = reads.shape[1] + refs.shape[1]
            assert np.all(f[:, 0] == f[:, 2]) and np.all(f[:, 1] == L - 1) and np.all(f[:, 3] == L - 1)
            bad = used_region_equal(a, b, f[:, 0], oa, ob, ostart)
            assert bad.size == 0, f"{kernel} {label} opt={opt} sc={sc}: pairs {bad[:5]} start ref={f[bad[:5],0]} oracle={ostart[bad[:5]]}"


def test_policies_really_differ():
    """Sanity: the two traceback policies are not the same function."""
    reads, refs = synth.uniform_batch(320, 64, 96, independent=True, seed=synth.BASE_SEED + 5)
    a0, b0, s0, _ = ora.align(ora.SW, ora.POLICY_DEFAULT_OCL, reads, refs, (5, -4, -1, -7))
    a1, b1, s1, _ = ora.align(ora.SW, ora.POLICY_SIMD, reads, refs, (5, -4, -1, -7))
    assert used_region_equal(a0, b0, s0, a1, b1, s1).size > 0
