import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from oracle import binding as ora
from versalignlib_b200 import capi, synth
from tests.helpers import used_region_equal
with capi.CudaContext(devices=[0]) as ctx:
    for name, (r, f) in {"c1": synth.uniform_batch(1000, 100, 150, p_sub=0.10, seed=1), "rand": synth.uniform_batch(500, 64, 96, independent=True, seed=5),
                         "tiny": synth.uniform_batch(65, 17, 9, p_sub=0.3, seed=7)}.items():
        for sc in [(2, -1, -3, -3), (3, -2, -1, -4), (1, 0, -7, -1)]:
            oa, ob, ostart, oend = ora.align(0, 1, r, f, sc)
            a, b, start, end = ctx.align_flat(0, 1, r, f, sc)
            bad = used_region_equal(a, b, start, oa, ob, ostart)
            print(name, sc, "end mismatch", int((end != oend).any(axis=1).sum()), "start mismatch", int((start != ostart).sum()), "bad", bad.size, "of", r.shape[0])
            if bad.size:
                i = bad[0]
                print(" pair", i, "start", start[i], ostart[i], "end", end[i], oend[i])
                L = a.shape[1]
                print("  got ", bytes(a[i, start[i]:L-1]).decode(errors='replace')[:80])
                print("  want", bytes(oa[i, ostart[i]:L-1]).decode(errors='replace')[:80])
                print("  gotR", bytes(b[i, start[i]:L-1]).decode(errors='replace')[:80])
                print("  wanR", bytes(ob[i, ostart[i]:L-1]).decode(errors='replace')[:80])
