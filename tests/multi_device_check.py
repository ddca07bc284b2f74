"""In-process multi-device check (run by hand on a box with >= 2 GPUs; not collected by pytest): one context
over all visible devices, the batch sharded inside the C ABI, results against the oracle.
usage: python tests/multi_device_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import binding as ora  # noqa: E402
from tests.helpers import used_region_equal  # noqa: E402
from versalignlib_b200 import capi, synth  # noqa: E402


def main():
    nd = capi.device_count()
    reads, refs, rl, fl = synth.mixed_batch(150_001, 20, 120, p_sub=0.1, q_indel=0.02, seed=9)
    with capi.CudaContext(devices=list(range(nd))) as ctx:
        for opt in (ora.SW, ora.NW):
            assert np.array_equal(ctx.score_flat(opt, reads, refs), ora.score(opt, reads, refs)), ("score", opt)
            a, b, start, end = ctx.align_flat(opt, 0, reads, refs)
            oa, ob, ostart, oend = ora.align(opt, 0, reads, refs)
            assert np.array_equal(start, ostart) and np.array_equal(end, oend), ("align", opt)
            assert used_region_equal(a, b, start, oa, ob, ostart).size == 0, ("strings", opt)
        pr, ro = synth.pack_batch(reads, rl)
        pf, fo = synth.pack_batch(refs, fl)
        tr, tf = np.ascontiguousarray(reads[:, :int(rl.max())]), np.ascontiguousarray(refs[:, :int(fl.max())])
        assert np.array_equal(ctx.score_packed(ora.SW, pr, ro, pf, fo), ora.score(ora.SW, tr, tf))
        scores, coords, coff, cigar = ctx.align_packed(ora.NW, 0, pr, ro, pf, fo)
        idx = np.arange(0, reads.shape[0], 37)
        want = synth.cigar_from_strings(*ora.align(ora.NW, 0, np.ascontiguousarray(tr[idx]), np.ascontiguousarray(tf[idx])))
        assert np.array_equal(coords[idx], want[0])
        got_runs = [cigar[coff[i]:coff[i + 1]].tolist() for i in idx]
        want_runs = [want[2][want[1][k]:want[1][k + 1]].tolist() for k in range(len(idx))]
        assert got_runs == want_runs
        print(f"multi-device ok: {nd} devices, {reads.shape[0]} mixed pairs, timings {ctx.timings()}")


if __name__ == "__main__":
    main()
