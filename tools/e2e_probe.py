"""End-to-end timing of one boundary at a time on the C2 workload (development aid; bench.py is the contract).
usage: python tools/e2e_probe.py [--pairs N] [--threads T] [--steps K] [--modes legacy,packed,packed_pinned]
  legacy         dlopen -> spawn -> compute_alignments, scattered char* in, new char[] out
  packed         va_cuda_align_packed from pageable numpy arrays
  packed_pinned  va_cuda_align_packed from page-locked arrays (read in place by the copy engines)"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from versalignlib_b200 import capi, synth  # noqa: E402
from versalignlib_b200.host import PluginHost  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--modes", default="legacy,packed,packed_pinned")
    ap.add_argument("--opt", type=int, default=1)
    a = ap.parse_args()
    threads = a.threads or (os.cpu_count() or 1)
    reads, refs = synth.uniform_batch(a.pairs, 150, 150, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 2)
    cells = a.pairs * 150 * 150
    out = {"pairs": a.pairs, "threads": threads, "malloc_tune": os.environ.get("VERSALIGN_CUDA_MALLOC_TUNE", "1")}
    sc = (2, -1, -3, -3)
    for mode in a.modes.split(","):
        times = []
        if mode == "legacy":
            h = PluginHost(capi.library_path(), 150, 150, sc, num_threads=threads, extra={"cuda_devices": 1}, verbosity=0)
            h.stage(reads, refs, scattered=True)
            for it in range(2 + a.steps):
                h.align_staged(a.opt, fetch=False)
                if it >= 2:
                    times.append(h.last_call_seconds)
                h.drop_alignments()
            ph = capi.plugin_timings()
            h.close()
        else:
            pr, ro = synth.pack_batch(reads)
            pf, fo = synth.pack_batch(refs)
            if mode == "packed_pinned":
                pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
                pr, ro, pf, fo = pin(pr), pin(ro), pin(pf), pin(fo)
            with capi.CudaContext(devices=[0], host_threads=threads) as ctx:
                keep = {}
                for it in range(2 + a.steps):
                    t0 = time.perf_counter()
                    ctx.align_packed(a.opt, 0, pr, ro, pf, fo, sc, out=keep)
                    if it >= 2:
                        times.append(time.perf_counter() - t0)
                ph = ctx.timings()
        sec = sum(times) / len(times)
        out[mode] = {"ms": round(sec * 1e3, 2), "min_ms": round(min(times) * 1e3, 2), "gcups": round(cells / sec / 1e9, 1),
                     "phases": {k: (round(v, 4) if isinstance(v, float) else v) for k, v in ph.items()}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
