"""Resident-path GCUPS of every mode on a chosen shape (development aid; bench.py is the contract).
usage: python tools/bench_modes.py [--pairs N] [--read L] [--ref L] [--policy P] [--mixed]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from versalignlib_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--read", type=int, default=150)
    ap.add_argument("--ref", type=int, default=150)
    ap.add_argument("--policy", type=int, default=0)
    ap.add_argument("--mixed", action="store_true", help="C3-style mixed lengths 100..ref, '\\0' padded")
    ap.add_argument("--scoring", type=int, nargs=4, default=[2, -1, -3, -3])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--modes", default="sw_score,nw_score,sw_align,nw_align")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    if a.mixed:
        reads, refs, rl, fl = synth.mixed_batch(a.pairs, 100, a.ref, p_sub=0.1, seed=synth.BASE_SEED + 3)
        cells = float((rl.astype(np.int64) * fl).sum())
    else:
        reads, refs = synth.uniform_batch(a.pairs, a.read, a.ref, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 2)
        cells = float(a.pairs) * a.read * a.ref
    n, L = reads.shape[0], reads.shape[1] + refs.shape[1]
    dr, df = torch.from_numpy(reads).to(dev), torch.from_numpy(refs).to(dev)
    ds = torch.zeros(n, dtype=torch.int16, device=dev)
    da = torch.empty((n, L), dtype=torch.uint8, device=dev)
    db = torch.empty((n, L), dtype=torch.uint8, device=dev)
    dst = torch.empty(n, dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    sc = tuple(a.scoring)
    out = {}
    with capi.CudaContext(devices=[0]) as ctx:
        for mode in a.modes.split(","):
            opt = 0 if mode.startswith("sw") else 1
            align = mode.endswith("align")

            def step():
                if align:
                    ctx.align_device(opt, a.policy, dr, df, da, db, dst, None, sc, stream=stream)
                else:
                    ctx.score_device(opt, dr, df, ds, sc, stream=stream)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            # per-kernel split: a separate, profiled pass (its sub-chunks run back to back, not overlapped)
            ctx.set_profiling(True)
            km = [0.0, 0.0, 0.0]
            for _ in range(a.steps):
                step()
                km = [x + y for x, y in zip(km, ctx.kernel_ms())]
            torch.cuda.synchronize()
            ctx.set_profiling(False)
            ms = e0.elapsed_time(e1) / a.steps
            out[mode] = {"ms": round(ms, 3), "gcups": round(cells / ms / 1e6, 1),
                         "prep_ms": round(km[0] / a.steps, 3), "fill_ms": round(km[1] / a.steps, 3),
                         "tb_ms": round(km[2] / a.steps, 3), "fill_gcups": round(cells / (km[1] / a.steps) / 1e6, 1)}
    print(json.dumps({"pairs": n, "read": reads.shape[1], "ref": refs.shape[1], "mixed": a.mixed, "scoring": sc,
                      "policy": a.policy, "modes": out}, indent=1))


if __name__ == "__main__":
    main()
