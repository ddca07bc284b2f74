"""Parity at the sizes BASELINE.json names (run by hand / by the round's evidence run on a GPU box; too long for the
pytest suite).  Every result of the CUDA path is compared with the OpenMP oracle -- not a sample:

  C2   1 M x (150 vs 150) NW compute_alignments: start, end cell, every byte of both gapped strings
  C3   10 M DISTINCT mixed-length pairs (100..250) SW score, in slices generated on the device
  C4   10 kbp x 12 kbp: SW scores of `--c4-score` pairs, SW and NW alignments of `--c4-align` pairs (the packed
       intra-task kernels + the warp-per-pair traceback)

usage: python tests/parity_at_scale.py [--out profiles/r2_parity_at_scale.json] [--quick]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import binding as ora  # noqa: E402
from tests.helpers import used_region_equal  # noqa: E402
from versalignlib_b200 import capi, synth  # noqa: E402

SCORING = (2, -1, -3, -3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join("gpurun_out", "parity_at_scale.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--c4-score", type=int, default=96)
    ap.add_argument("--c4-align", type=int, default=24)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    out = {"oracle": "oracle/va_oracle.c (OpenMP, all host cores)", "host_cores": os.cpu_count()}
    with capi.CudaContext(devices=[0]) as ctx:
        # ---- C2: the full batch
        n = 100_000 if a.quick else 1_000_000
        reads, refs = synth.uniform_batch(n, 150, 150, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 2)
        t0 = time.perf_counter()
        ga, gb, gstart, gend = ctx.align_flat(ora.NW, 0, reads, refs)
        t_gpu = time.perf_counter() - t0
        bad = 0
        t0 = time.perf_counter()
        for lo in range(0, n, 250_000):
            hi = min(n, lo + 250_000)
            oa, ob, ostart, oend = ora.align(ora.NW, 0, reads[lo:hi], refs[lo:hi], SCORING)
            bad += int((gstart[lo:hi] != ostart).sum()) + int((gend[lo:hi] != oend).any(axis=1).sum())
            bad += int(used_region_equal(ga[lo:hi], gb[lo:hi], gstart[lo:hi], oa, ob, ostart).size)
        out["C2_nw_align_1M_150x150"] = {"pairs": n, "compared": "start, end cell, every byte of both strings, all pairs",
                                         "mismatches": bad, "cuda_seconds_flat_api": round(t_gpu, 3), "oracle_seconds": round(time.perf_counter() - t0, 1)}
        del ga, gb, reads, refs

        # ---- C3: 10 M distinct pairs in slices
        total, slice_pairs = (1_000_000, 500_000) if a.quick else (10_000_000, 1_000_000)
        bad, cells, t_or = 0, 0.0, 0.0
        stream = torch.cuda.current_stream().cuda_stream
        for k, lo in enumerate(range(0, total, slice_pairs)):
            reads, refs, rl, fl = synth.mixed_batch_torch(slice_pairs, 100, 250, 0.10, synth.BASE_SEED + 3 + 7919 * k, dev)
            d_s = torch.zeros(slice_pairs, dtype=torch.int16, device=dev)
            ctx.score_device(ora.SW, reads, refs, d_s, SCORING, stream=stream)
            torch.cuda.synchronize()
            cells += float((rl.long() * fl.long()).sum().item())
            t0 = time.perf_counter()
            want = ora.score(ora.SW, reads.cpu().numpy(), refs.cpu().numpy(), SCORING)
            t_or += time.perf_counter() - t0
            bad += int((d_s.cpu().numpy() != want).sum())
            del reads, refs, d_s
        out["C3_sw_score_10M_mixed_100_250"] = {"pairs": total, "distinct": True, "cells": cells, "compared": "every score",
                                                "mismatches": bad, "oracle_seconds": round(t_or, 1)}

        # ---- C4: long pairs
        ns, na = (8, 4) if a.quick else (a.c4_score, a.c4_align)
        reads, refs = synth.uniform_batch(max(ns, na), 10_000, 12_000, p_sub=0.10, q_indel=0.03, seed=synth.BASE_SEED + 4)
        t0 = time.perf_counter()
        want = ora.score(ora.SW, reads[:ns], refs[:ns], SCORING)
        got = ctx.score_flat(ora.SW, np.ascontiguousarray(reads[:ns]), np.ascontiguousarray(refs[:ns]), SCORING)
        c4 = {"score_pairs": ns, "score_mismatches": int((got != want).sum())}
        r, f = np.ascontiguousarray(reads[:na]), np.ascontiguousarray(refs[:na])
        for name, opt, pols in (("sw_align", ora.SW, (0,)), ("nw_align", ora.NW, (0, 1))):
            for pol in pols:
                oa, ob, ostart, oend = ora.align(opt, pol, r, f, SCORING)
                ga, gb, gstart, gend = ctx.align_flat(opt, pol, r, f, SCORING)
                bad = int((gstart != ostart).sum()) + int((gend != oend).any(axis=1).sum()) + int(used_region_equal(ga, gb, gstart, oa, ob, ostart).size)
                c4[f"{name}_policy{pol}"] = {"pairs": na, "mismatches": bad, "mean_alignment_columns": float((ga.shape[1] - 1 - gstart.astype(np.int64)).mean())}
        c4["seconds"] = round(time.perf_counter() - t0, 1)
        out["C4_10kbp_x_12kbp"] = c4
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out, indent=1))
    total_bad = out["C2_nw_align_1M_150x150"]["mismatches"] + out["C3_sw_score_10M_mixed_100_250"]["mismatches"] + c4["score_mismatches"] + sum(
        v["mismatches"] for v in c4.values() if isinstance(v, dict))
    sys.exit(1 if total_bad else 0)


if __name__ == "__main__":
    main()
