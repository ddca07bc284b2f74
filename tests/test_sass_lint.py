"""Build-time guard (no GPU): the row loops of the packed fill kernels must not contain the LOP3 storm
ptxas produces when it parks the direction predicates in a register (tools/check_sass.py)."""
import os
import shutil
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_row_loops_are_clean():
    import check_sass
    from versalignlib_b200 import build
    lib = build.build_cuda()
    report = check_sass.loops_of(lib)
    assert len(report) >= 16, "expected every packed instantiation to have a row loop"
    for name, _, c in report:
        dpx = sum(c[k] for k in check_sass.DPX)
        assert c["LOP3"] <= dpx, (name, dict(c))
