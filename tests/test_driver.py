"""The driver / benchmark harness (versalignlib_b200/driver.py): any plug-in next to any other through the
reference's own boundary, a parity verdict and GCUPS in one JSON object."""
import json
import os

import numpy as np
import pytest

from oracle import binding as ora
from versalignlib_b200 import driver, synth


def _run(capsys, argv):
    rc = driver.main(argv)
    return rc, json.loads(capsys.readouterr().out.strip().splitlines()[-1])


@pytest.mark.skipif(ora.ref_lib("Default") is None or ora.ref_lib("AVX") is None, reason="reference kernels not built")
def test_reference_kernels_side_by_side(capsys):
    rc, out = _run(capsys, ["--kernel", ora.ref_lib("AVX"), "--threads", "1", "--compare", ora.ref_lib("SSE"),
                            "--compare-threads", "1", "--mode", "nw_score", "--synthetic", "1600,64,96", "--reps", "1"])
    assert rc == 0 and out["compare"]["parity"]["mismatches"] == 0 and out["gcups"] > 0


@pytest.mark.skipif(ora.ref_lib("Default") is None, reason="reference kernels not built")
def test_fasta_input(capsys, tmp_path):
    reads, refs, rl, fl = synth.mixed_batch(320, 20, 90, seed=5)
    for name, arr, lens in (("reads.fa", reads, rl), ("refs.fa", refs, fl)):
        with open(tmp_path / name, "wb") as f:
            for i in range(arr.shape[0]):
                f.write(b">s%d\n%s\n" % (i, arr[i, :lens[i]].tobytes()))
    rc, out = _run(capsys, ["--kernel", ora.ref_lib("Default"), "--compare", ora.ref_lib("Default"), "--mode", "sw_align",
                            "--reads", str(tmp_path / "reads.fa"), "--refs", str(tmp_path / "refs.fa"), "--reps", "1"])
    assert rc == 0 and out["pairs"] == 320 and out["read_length"] == int(rl.max()) and out["ref_length"] == int(fl.max())
    assert out["compare"]["parity"]["mismatches"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode,policy,ref,threads", [("nw_align", 0, "Default", 0), ("sw_align", 1, "AVX", 1),
                                                      ("sw_score", 0, "AVX", 1), ("nw_score", 0, "SSE", 1)])
def test_cuda_kernel_against_reference_kernels(capsys, mode, policy, ref, threads):
    if ora.ref_lib(ref) is None:
        pytest.skip("reference kernels not built")
    argv = ["--mode", mode, "--policy", str(policy), "--compare", ora.ref_lib(ref), "--synthetic", "3200,100,150", "--reps", "1"]
    if threads:
        argv += ["--compare-threads", str(threads)]
    rc, out = _run(capsys, argv)
    assert rc == 0 and out["compare"]["parity"]["mismatches"] == 0, out
