"""Device-resident C2 step (NW compute_alignments, 150 x 150) timed alone (development aid; bench.py is the contract).
usage: python tools/resident_probe.py [--pairs N] [--steps K]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from versalignlib_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--opt", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
reads, refs = synth.uniform_batch(a.pairs, 150, 150, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 2)
dr, df = torch.from_numpy(reads).to(dev), torch.from_numpy(refs).to(dev)
n, L = a.pairs, 300
da = torch.empty((n, L), dtype=torch.uint8, device=dev)
db = torch.empty((n, L), dtype=torch.uint8, device=dev)
dst = torch.empty(n, dtype=torch.int16, device=dev)
de = torch.empty((n, 2), dtype=torch.int16, device=dev)
stream = torch.cuda.current_stream().cuda_stream
with capi.CudaContext(devices=[0]) as ctx:
    for _ in range(3):
        ctx.align_device(a.opt, 0, dr, df, da, db, dst, de, stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ctx.align_device(a.opt, 0, dr, df, da, db, dst, de, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    ctx.set_profiling(True)
    ctx.align_device(a.opt, 0, dr, df, da, db, dst, de, stream=stream)
    ph = ctx.kernel_ms()
    ctx.set_profiling(False)
    print(f"pairs {n} ms/step {ms:.3f} GCUPS {n * 22500 / ms / 1e6:.0f} checksum {int(dst.to(torch.int64).sum())} "
          f"prep/fill/traceback ms {ph[0]:.3f} {ph[1]:.3f} {ph[2]:.3f}")
