"""Golden fixtures = outputs of the reference's own kernels (tests/golden/make_golden.py).
CPU: the oracle reproduces them.  GPU: the CUDA path reproduces them, through the C ABI."""
import glob
import os

import numpy as np
import pytest

from oracle import binding as ora
from tests.helpers import used_region_equal

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def _cases():
    for path in GOLDEN:
        z = np.load(path)
        for pi, sc in enumerate(z["params"]):
            yield os.path.basename(path)[:-4], z, pi, tuple(int(x) for x in sc)


def test_fixtures_present():
    assert len(GOLDEN) >= 6


def test_oracle_reproduces_reference_outputs():
    for label, z, pi, sc in _cases():
        reads, refs = z["reads"], z["refs"]
        for opt, m in ((0, "sw"), (1, "nw")):
            assert np.array_equal(ora.score(opt, reads, refs, sc), z[f"p{pi}_{m}_score"]), (label, sc, m)
            for pol, tag in ((ora.POLICY_DEFAULT_OCL, "default"), (ora.POLICY_SIMD, "simd")):
                a, b, start, _ = ora.align(opt, pol, reads, refs, sc)
                bad = used_region_equal(a, b, start, z[f"p{pi}_{m}_aln_read_{tag}"], z[f"p{pi}_{m}_aln_ref_{tag}"],
                                        z[f"p{pi}_{m}_start_{tag}"])
                assert bad.size == 0, (label, sc, m, tag, bad[:5])


@pytest.mark.gpu
def test_cuda_reproduces_reference_outputs():
    from versalignlib_b200 import capi
    with capi.CudaContext(devices=[0]) as ctx:
        for label, z, pi, sc in _cases():
            reads, refs = z["reads"], z["refs"]
            for opt, m in ((0, "sw"), (1, "nw")):
                assert np.array_equal(ctx.score_flat(opt, reads, refs, sc), z[f"p{pi}_{m}_score"]), (label, sc, m)
                for pol, tag in ((0, "default"), (1, "simd")):
                    a, b, start, _ = ctx.align_flat(opt, pol, reads, refs, sc)
                    bad = used_region_equal(a, b, start, z[f"p{pi}_{m}_aln_read_{tag}"], z[f"p{pi}_{m}_aln_ref_{tag}"],
                                            z[f"p{pi}_{m}_start_{tag}"])
                    assert bad.size == 0, (label, sc, m, tag, bad[:5])
