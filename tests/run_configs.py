"""Evidence run for the BASELINE.json configs other than the bench headline (C2): one JSON object with
GCUPS and parity verdicts, written to the path given (default gpurun_out/configs.json).
  C1  SW score, 100k x (100 vs 150): through the plug-in boundary next to the reference's SSE (1 thread)
      and Default (all cores) kernels -- parity verdict + speed-up (the reference driver's own case)
  C3  SW score, mixed 100-250: one GPU's share of the 10M batch (1.25M pairs) resident, and the whole
      10M batch end to end through va_cuda_score_packed on this one GPU
  C4  SW, 10k x (10 kbp vs 12 kbp): scores resident (intra-task kernel), a 1250-pair share (8-GPU split),
      and alignments of a declared subset (8 pairs) checked against the oracle
usage: python tests/run_configs.py [out.json] [--quick]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import binding as ora  # noqa: E402
from versalignlib_b200 import capi, driver, synth  # noqa: E402


def resident_score(ctx, opt, reads, refs, cells, steps=3):
    dev = torch.device("cuda:0")
    dr, df = torch.from_numpy(reads).to(dev), torch.from_numpy(refs).to(dev)
    ds = torch.zeros(reads.shape[0], dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        ctx.score_device(opt, dr, df, ds, stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ctx.score_device(opt, dr, df, ds, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms": round(ms, 3), "gcups": round(cells / ms / 1e6, 1)}, ds.cpu().numpy()


def main():
    out_path = next((a for a in sys.argv[1:] if not a.startswith("--")), "gpurun_out/configs.json")
    quick = "--quick" in sys.argv
    out = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count()}
    import io
    from contextlib import redirect_stdout

    # ---- C1 through the reference's boundary, next to the reference's kernels
    c1 = {}
    for ref, threads in (("SSE", 1), ("Default", os.cpu_count() or 1)):
        if ora.ref_lib(ref) is None:
            continue
        buf = io.StringIO()
        with redirect_stdout(buf):
            driver.main(["--mode", "sw_score", "--synthetic", "100000,100,150", "--compare", ora.ref_lib(ref),
                         "--compare-threads", str(threads), "--reps", "3"])
        d = json.loads(buf.getvalue().strip().splitlines()[-1])
        c1[f"vs_{ref}"] = {"gcups_e2e": round(d["gcups"], 1), "seconds": d["seconds_median"],
                           "reference_gcups": round(d["compare"]["gcups"], 3), "reference_threads": threads,
                           "speedup": round(d["compare"]["speedup"], 1), "parity": d["compare"]["parity"]}
    out["C1_sw_score_100k_100x150_plugin_boundary"] = c1

    with capi.CudaContext(devices=[0]) as ctx:
        # ---- C3
        n3 = 200_000 if quick else 1_250_000
        reads, refs, rl, fl = synth.mixed_batch(n3, 100, 250, p_sub=0.1, seed=synth.BASE_SEED + 3)
        cells = float((rl.astype(np.int64) * fl).sum())
        res, got = resident_score(ctx, ora.SW, reads, refs, cells)
        idx = np.random.default_rng(3).choice(n3, 4000, replace=False)
        res["oracle_sample_mismatches"] = int((got[idx] != ora.score(ora.SW, np.ascontiguousarray(reads[idx]), np.ascontiguousarray(refs[idx]))).sum())
        res["pairs"] = n3
        res["cells_counted"] = "true rows x cols per pair"
        out["C3_sw_score_mixed_100_250_resident_share"] = res
        reps = 1 if quick else 8  # 8 x 1.25M = the 10M batch
        pr, ro = synth.pack_batch(reads, rl)
        pf, fo = synth.pack_batch(refs, fl)
        big_r = np.tile(pr, reps); big_f = np.tile(pf, reps)
        big_ro = np.concatenate([[0], np.cumsum(np.tile(np.diff(ro), reps))]).astype(np.int64)
        big_fo = np.concatenate([[0], np.cumsum(np.tile(np.diff(fo), reps))]).astype(np.int64)
        ctx.score_packed(ora.SW, big_r, big_ro, big_f, big_fo)  # warm-up: buffers
        t0 = time.perf_counter()
        sc = ctx.score_packed(ora.SW, big_r, big_ro, big_f, big_fo)
        dt = time.perf_counter() - t0
        out["C3_sw_score_mixed_10M_e2e_packed_one_gpu"] = {
            "pairs": int(n3 * reps), "seconds": round(dt, 4), "gcups": round(cells * reps / dt / 1e9, 1),
            "agrees_with_resident": bool(np.array_equal(sc[:n3], got)), "phases": ctx.timings()}
        del big_r, big_f, reads, refs

        # ---- C4
        for tag, n4 in (("C4_sw_score_10kbp_x_12kbp_resident", 1000 if quick else 10_000), ("C4_sw_score_share_of_8_gpus", 1250)):
            reads, refs = synth.uniform_batch(n4, 10_000, 12_000, p_sub=0.10, q_indel=0.03, seed=synth.BASE_SEED + 4)
            res, got = resident_score(ctx, ora.SW, reads, refs, float(n4) * 10_000 * 12_000, steps=2)
            k = 6
            res["oracle_sample_mismatches"] = int((got[:k] != ora.score(ora.SW, np.ascontiguousarray(reads[:k]), np.ascontiguousarray(refs[:k]))).sum())
            res["pairs"] = n4
            out[tag] = res
        n_aln, n_chk = (64, 4) if quick else (592, 8)  # 592 pairs = 296 pair-of-pairs = one CTA wave of the intra-task kernel at 8 warps
        sub_r, sub_f = np.ascontiguousarray(reads[:n_aln]), np.ascontiguousarray(refs[:n_aln])
        ctx.align_flat(ora.SW, 0, sub_r, sub_f)  # warm-up: workspace allocation
        t0 = time.perf_counter()
        a, b, start, end = ctx.align_flat(ora.SW, 0, sub_r, sub_f)
        dt = time.perf_counter() - t0
        oa, ob, ostart, oend = ora.align(ora.SW, 0, np.ascontiguousarray(sub_r[:n_chk]), np.ascontiguousarray(sub_f[:n_chk]))
        L = a.shape[1]
        col = np.arange(L)[None, :]
        used = (col >= np.clip(ostart.astype(np.int64), 0, L)[:, None]) & (col < L - 1)
        bad = (start[:n_chk] != ostart) | (end[:n_chk] != oend).any(axis=1) | ((a[:n_chk] != oa) & used).any(axis=1) | ((b[:n_chk] != ob) & used).any(axis=1)
        out["C4_sw_align_declared_subset"] = {"pairs": n_aln, "seconds_flat_api": round(dt, 3), "gcups_flat_api": round(n_aln * 1.2e8 / dt / 1e9, 1),
                                              "oracle_checked_pairs": n_chk, "mismatches_vs_oracle": int(bad.sum()),
                                              "kernel": "packed intra-task fill (va_intra.cu) + warp-per-pair traceback"}
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
