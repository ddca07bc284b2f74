#!/usr/bin/env python
"""bench.py -- GCUPS of the batched DP hot path on B200, one process per GPU.

Workload (BASELINE.json configs[1], "C2"): the library's "Needleman-Wunsch" compute_alignments
(score + traceback, gapped strings out) on 1 M synthetic 150 bp read/ref pairs PER GPU
(weak scaling: pairs are independent, no collective on the data path).

  value        whole-job GCUPS with the raw batch already resident in HBM: prep + fill +
               traceback kernels through va_cuda_align_device on torch's current stream,
               CUDA events, max over ranks
  e2e          the same metric through the reference-facing plug-in boundary
               (dlopen -> spawn_alignment_kernel -> AlignmentKernel::compute_alignments) with
               scattered HOST buffers in and new char[] blocks out; H2D/D2H inside the timed region
  e2e_packed   the batch-friendly entry point beside it (va_cuda_align_packed): offset-addressed
               sequences in PAGE-LOCKED host memory in (read in place by the copy engines), scores +
               coordinates + BAM-style CIGARs out
  e2e_inprocess  (N > 1) one compute_alignments call on rank 0 with cuda_devices = N: the in-process shard
               over the N GPUs that BASELINE.json's north_star describes, on N x the pairs
  roofline     the fill kernel against the measured integer-pipe peak (VIADDMNMX.S16x2 lane-ops/s
               divided by 3 ops per cell), plus its HBM view
  modes        resident GCUPS of every function (both pointer policies; the 32-bit general kernel through the affine-gap
               variant) on the same shape
  configs      the other BASELINE configs at these N GPUs: C1 through the plug-in boundary next to the
               reference's SSE / AVX / Default kernels, C3 (10 M distinct mixed-length pairs, strong-scaled over
               the ranks, cells = true rows x cols), C4 (10 k long pairs split over the ranks); each with resident
               and end-to-end GCUPS, the roofline fraction and an oracle-sample mismatch count
  cpu_baseline / --impl reference
               the reference's own CPU kernel (oracle/_ref/libDefaultKernel.so, OpenMP, all host
               cores) through the same plug-in boundary on a bounded sample of the same workload

A "step" = one pass of the hot path over one batch.  Cells are counted as rows x cols actually
required (150 x 150 per pair here; nothing is padded in this workload).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))

import numpy as np  # noqa: E402

READ_LEN = 150
REF_LEN = 150
PAIRS_PER_GPU = 1_000_000
SCORING = (2, -1, -3, -3)
SW, NW = 0, 1
POLICY_DEFAULT_OCL, POLICY_SIMD = 0, 1
CELLS_PER_PAIR = READ_LEN * REF_LEN
WORKLOAD = "C2: NW compute_alignments (score+traceback), 1M x (150bp read vs 150bp ref) per GPU, p_sub=0.08 q_indel=0.02"
REFERENCE_SAMPLE_PAIRS = 48_000


def config_dict(pairs_per_gpu: int, world: int) -> dict:
    """The same object in both arms (ours and --impl reference)."""
    return {"workload": WORKLOAD, "pairs_per_gpu": pairs_per_gpu, "read_length": READ_LEN, "ref_length": REF_LEN,
            "scoring": list(SCORING), "traceback_policy": "DEFAULT_OCL",
            "parallelism": f"dp{world} (independent pairs, no collective)",
            "l2": "inputs + direction matrix (5.9 GB/step) exceed the 126 MB L2; no explicit flush"}


def make_batch(n: int, seed_offset: int = 0):
    from versalignlib_b200 import synth
    return synth.uniform_batch(n, READ_LEN, REF_LEN, p_sub=0.08, q_indel=0.02, seed=synth.BASE_SEED + 2 + 1000 * seed_offset)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines: list[str] = []
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def tune_host_malloc() -> str:
    """Host-application side: keep freed heap pages inside glibc's arenas between calls (the CUDA plug-in asks
    for the same thing itself when it is spawned; the flag makes the reference arm's host do it too)."""
    import ctypes
    try:
        libc = ctypes.CDLL("libc.so.6")
        M_TRIM_THRESHOLD, M_TOP_PAD, M_MMAP_THRESHOLD = -1, -2, -3
        libc.mallopt(M_TRIM_THRESHOLD, 1 << 30)
        libc.mallopt(M_TOP_PAD, 64 << 20)
        libc.mallopt(M_MMAP_THRESHOLD, 1 << 30)
        return "mallopt(M_TRIM_THRESHOLD=1GiB, M_TOP_PAD=64MiB, M_MMAP_THRESHOLD=1GiB)"
    except Exception as e:  # pragma: no cover
        return f"default ({e})"


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "MEASURED_PEAKS.json"
        return d
    return {"hbm_gbs": 6650.0, "_source": "fallback (B200_PROFILING.md)"}


def limiter_of(ph: dict, ms_total: float) -> str:
    """Which resource the end-to-end call waits on, from its phase counters."""
    if not ph:
        return "unknown"
    host = (ph.get("gather_s", 0) + ph.get("scatter_s", 0)) * 1e3
    kern = ph.get("kernel_ms", 0.0)
    pcie = max(ph.get("h2d_bytes", 0), ph.get("d2h_bytes", 0)) / 50e9 * 1e3 / max(1, ph.get("devices", 1))  # ~50 GB/s per direction per GPU
    parts = {"host staging (gather + scatter on the CPU threads)": host, "device kernels": kern, "PCIe copies": pcie}
    name = max(parts, key=parts.get)
    return f"{name}: {parts[name]:.1f} ms of {ms_total:.1f} ms (host {host:.1f}, kernels {kern:.1f}, PCIe >= {pcie:.1f})"


def reference_arm(steps: int, warmup: int, sample_pairs: int, threads: int):
    """The reference's own CPU kernel through its plug-in boundary on a bounded sample."""
    from oracle import binding as ora
    from versalignlib_b200.host import PluginHost
    lib = ora.ref_lib("Default")
    kind = "reference"
    reads, refs = make_batch(sample_pairs, seed_offset=7)
    times = []
    if lib is not None:
        with PluginHost(lib, READ_LEN, REF_LEN, SCORING, num_threads=threads, verbosity=0) as h:
            h.stage(reads, refs, scattered=True)
            for it in range(warmup + steps):
                h.align_staged(NW, fetch=False)
                if it >= warmup:
                    times.append(h.last_call_seconds)
                h.drop_alignments()
        sample = f"{sample_pairs} pairs of the same workload per step, libDefaultKernel.so (reference, OpenMP num_threads={threads})"
    else:  # reference tree was not available at build time: time the oracle port instead
        kind = "port"
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            ora.align(NW, ora.POLICY_DEFAULT_OCL, reads, refs, SCORING, threads=threads)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        sample = f"{sample_pairs} pairs of the same workload per step, oracle port (OpenMP threads={threads})"
    sec = sum(times) / len(times)
    gcups = sample_pairs * CELLS_PER_PAIR / sec / 1e9
    return {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": kind, "sample": sample, "ms_per_step": sec * 1e3}


def run_reference(args):
    if RANK != 0:
        return
    threads = os.cpu_count() or 1
    # ~2-4 GCUPS expected: 48k pairs x 22.5k cells ~ 1.1 Gcells -> a fraction of a second per step
    cb = reference_arm(args.steps, args.warmup, REFERENCE_SAMPLE_PAIRS, threads)
    line = {
        "impl": "reference", "metric": "GCUPS", "value": cb["value"], "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": config_dict(args.pairs, max(args.gpus, 1)),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU reference arm on rank 0 only: every step is a bounded sample of the workload (GCUPS is per cell, so the "
                "sample size does not enter the ratio); host_malloc: " + args.host_malloc,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# the other BASELINE configs (driver-visible evidence; each bounded to a few seconds)
# ------------------------------------------------------------------------------------------------

def timed_resident(torch, fn, steps: int):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def config_c1(torch, ctx, dev, roof_gcups):
    """C1: SW score_alignments, 100k x (100 vs 150): through the plug-in boundary next to the reference's kernels."""
    from oracle import binding as ora
    from versalignlib_b200 import capi, driver, synth
    n = 100_000
    reads, refs = synth.uniform_batch(n, 100, 150, p_sub=0.10, seed=synth.BASE_SEED + 1)
    cells = float(n) * 100 * 150
    out = {"workload": "C1: SW score_alignments, 100k x (100bp vs 150bp), through dlopen + the virtual call", "pairs": n}
    ours, t_ours = driver.run(capi.library_path(), SW, False, reads, refs, SCORING, os.cpu_count() or 1, 3,
                              {"cuda_devices": 1, "cuda_device_first": dev.index})
    out["e2e_gcups"] = cells / statistics.median(t_ours) / 1e9
    cpu = {}
    for name, threads in (("SSE", 1), ("AVX", 1), ("Default", os.cpu_count() or 1)):
        lib = ora.ref_lib(name)
        if lib is None:
            continue
        res, t = driver.run(lib, SW, False, reads, refs, SCORING, threads, 1, {})
        got = res[0]
        # Default stores only the low byte of a score (DefaultKernel.cpp:137): compare that byte
        same = np.array_equal(got & 0xFF, ours[0] & 0xFF) if name == "Default" else np.array_equal(got, ours[0])
        cpu[name] = {"gcups": cells / statistics.median(t) / 1e9, "threads": threads, "mismatches_vs_cuda": 0 if same else int((got != ours[0]).sum())}
    out["reference_kernels"] = cpu
    d_r, d_f = torch.from_numpy(reads).to(dev), torch.from_numpy(refs).to(dev)
    d_s = torch.zeros(n, dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ms = timed_resident(torch, lambda: ctx.score_device(SW, d_r, d_f, d_s, SCORING, stream=stream), 5)
    out["resident_gcups"] = cells / ms / 1e6
    out["roofline_frac"] = out["resident_gcups"] / roof_gcups
    out["oracle_mismatches"] = int((d_s.cpu().numpy() != ora.score(ora.SW, reads, refs, SCORING)).sum())
    return out


def config_c3(torch, ctx, dev, roof_gcups, max_over_ranks, barrier, quick: bool):
    """C3: SW score on 10 M DISTINCT mixed-length pairs (100..250), strong-scaled: this rank owns 10M / WORLD."""
    from oracle import binding as ora
    from versalignlib_b200 import synth
    total = 1_000_000 if quick else 10_000_000
    n = total // WORLD
    reads, refs, rl, fl = synth.mixed_batch_torch(n, 100, 250, 0.10, synth.BASE_SEED + 3 + 7919 * RANK, dev)
    cells_local = float((rl.long() * fl.long()).sum().item())
    cells_t = torch.tensor([cells_local], dtype=torch.float64, device=dev)
    if WORLD > 1:
        import torch.distributed as dist
        dist.all_reduce(cells_t)
    cells = float(cells_t.item())
    d_s = torch.zeros(n, dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    barrier()
    ms = max_over_ranks(timed_resident(torch, lambda: ctx.score_device(SW, reads, refs, d_s, SCORING, stream=stream), 3))
    out = {"workload": f"C3: SW score, {total} distinct pairs, read ~U{{100..250}}, ref ~U{{read..250}}, p_sub=0.10, length-bucketed on the device; "
                       f"strong scaling: {n} pairs per rank", "pairs": total, "cells_counted": "true rows x cols per pair",
           "resident_ms": ms, "resident_gcups": cells / ms / 1e6, "scaling": "strong"}
    out["roofline_frac_per_gpu"] = out["resident_gcups"] / WORLD / roof_gcups
    # oracle on a sample of this rank's pairs
    g = torch.Generator(device="cpu"); g.manual_seed(3 + RANK)
    idx = torch.randperm(n, generator=g)[:4000].to(dev)
    got = d_s[idx].cpu().numpy()
    want = ora.score(ora.SW, np.ascontiguousarray(reads[idx].cpu().numpy()), np.ascontiguousarray(refs[idx].cpu().numpy()), SCORING)
    bad = torch.tensor([float((got != want).sum())], dtype=torch.float64, device=dev)
    if WORLD > 1:
        dist.all_reduce(bad)
    out["oracle_sample"] = {"pairs": 4000 * WORLD, "mismatches": int(bad.item())}
    # end to end: the batch-friendly entry point from page-locked host arrays (the legacy boundary cannot express
    # per-pair lengths without padding every sequence to 250)
    pr, ro = synth.pack_batch_torch(reads, rl)
    pf, fo = synth.pack_batch_torch(refs, fl)
    resident_scores = d_s.cpu().numpy().copy()
    del reads, refs, d_s
    torch.cuda.empty_cache()
    ctx.score_packed(SW, pr, ro, pf, fo, SCORING)  # warm-up: workspace
    times = []
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        sc = ctx.score_packed(SW, pr, ro, pf, fo, SCORING)
        times.append(max_over_ranks(time.perf_counter() - t0))
    sec = min(times)
    ph = ctx.timings()
    out["e2e_gcups"] = cells / sec / 1e9
    out["e2e_ms"] = sec * 1e3
    out["e2e_api"] = "va_cuda_score_packed, offset-addressed sequences in page-locked host memory"
    out["e2e_agrees_with_resident"] = bool(np.array_equal(sc, resident_scores))
    out["e2e_limiter"] = limiter_of(ph, sec * 1e3)
    return out


def config_c4(torch, ctx, dev, roof_gcups, max_over_ranks, barrier, quick: bool):
    """C4: SW on 10 k long pairs (10 kbp vs 12 kbp), split over the ranks: scores on all, alignments on a declared subset."""
    from oracle import binding as ora
    from versalignlib_b200 import synth
    total = 1_000 if quick else 10_000
    n = total // WORLD
    reads, refs = synth.uniform_batch(n, 10_000, 12_000, p_sub=0.10, q_indel=0.03, seed=synth.BASE_SEED + 4 + 7919 * RANK)
    cells = float(total) * 10_000 * 12_000
    d_r, d_f = torch.from_numpy(reads).to(dev), torch.from_numpy(refs).to(dev)
    d_s = torch.zeros(n, dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    barrier()
    ms = max_over_ranks(timed_resident(torch, lambda: ctx.score_device(SW, d_r, d_f, d_s, SCORING, stream=stream), 2))
    out = {"workload": f"C4: SW, {total} pairs of 10 kbp reads vs 12 kbp refs, p_sub=0.10 q_indel=0.03; {n} pairs per rank",
           "pairs": total, "score_resident_ms": ms, "score_resident_gcups": cells / ms / 1e6, "scaling": "strong"}
    out["score_roofline_frac_per_gpu"] = out["score_resident_gcups"] / WORLD / roof_gcups
    k = 4
    got = d_s[:k].cpu().numpy()
    want = ora.score(ora.SW, np.ascontiguousarray(reads[:k]), np.ascontiguousarray(refs[:k]), SCORING)
    out["score_oracle_sample"] = {"pairs": k * WORLD, "mismatches": int((got != want).sum())}
    # end to end (host buffers -> scores on the host)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    h_r, h_f = pin(reads), pin(refs)
    ctx.score_flat(SW, h_r, h_f, SCORING)
    barrier()
    t0 = time.perf_counter()
    sc = ctx.score_flat(SW, h_r, h_f, SCORING)
    sec = max_over_ranks(time.perf_counter() - t0)
    out["score_e2e_gcups"] = cells / sec / 1e9
    out["score_e2e_agrees_with_resident"] = bool(np.array_equal(sc, d_s.cpu().numpy()))
    # compute_alignments on a declared subset: the first n_aln pairs of every rank
    n_aln = min(n, 64 if quick else 1184)  # 1184 pairs = 592 pair-of-pairs = one CTA wave of the intra-task kernel at 4 warps per CTA
    L = 22_000
    sub_r, sub_f = d_r[:n_aln].contiguous(), d_f[:n_aln].contiguous()
    d_a = torch.empty((n_aln, L), dtype=torch.uint8, device=dev)
    d_b = torch.empty((n_aln, L), dtype=torch.uint8, device=dev)
    d_st = torch.empty(n_aln, dtype=torch.int16, device=dev)
    d_end = torch.empty((n_aln, 2), dtype=torch.int16, device=dev)
    ctx.set_profiling(True)
    fn = lambda: ctx.align_device(SW, POLICY_DEFAULT_OCL, sub_r, sub_f, d_a, d_b, d_st, d_end, SCORING, stream=stream)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    kms = ctx.kernel_ms()
    ctx.set_profiling(False)
    ams = max_over_ranks(e0.elapsed_time(e1))
    acells = float(n_aln) * WORLD * 10_000 * 12_000
    out["align_subset"] = {"pairs": n_aln * WORLD, "resident_ms": ams, "resident_gcups": acells / ams / 1e6,
                           "fill_gcups_this_rank": float(n_aln) * 1.2e8 / max(kms[1], 1e-6) / 1e6,
                           "ms_prep_fill_traceback_this_rank": [round(x, 3) for x in kms]}
    ka = 2
    oa, ob, ostart, oend = ora.align(ora.SW, 0, np.ascontiguousarray(reads[:ka]), np.ascontiguousarray(refs[:ka]), SCORING)
    a, b = d_a[:ka].cpu().numpy(), d_b[:ka].cpu().numpy()
    st, en = d_st[:ka].cpu().numpy(), d_end[:ka].cpu().numpy()
    col = np.arange(L)[None, :]
    used = (col >= np.clip(ostart.astype(np.int64), 0, L)[:, None]) & (col < L - 1)
    bad = (st != ostart) | (en != oend).any(axis=1) | ((a != oa) & used).any(axis=1) | ((b != ob) & used).any(axis=1)
    out["align_subset"]["oracle_sample"] = {"pairs": ka * WORLD, "mismatches": int(bad.sum())}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from versalignlib_b200 import capi, synth
    from versalignlib_b200.host import PluginHost

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    dev_index = LOCAL_RANK if torch.cuda.device_count() > LOCAL_RANK else 0
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    distributed = WORLD > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.pairs
    reads, refs = make_batch(n, seed_offset=RANK)
    L = READ_LEN + REF_LEN
    cells_per_step = n * CELLS_PER_PAIR

    ctx = capi.CudaContext(devices=[dev_index])
    stream = torch.cuda.current_stream().cuda_stream

    # ---------------- kernel-only: inputs resident in HBM ----------------
    d_reads = torch.from_numpy(reads).to(dev)
    d_refs = torch.from_numpy(refs).to(dev)
    d_a = torch.empty((n, L), dtype=torch.uint8, device=dev)
    d_b = torch.empty((n, L), dtype=torch.uint8, device=dev)
    d_start = torch.empty(n, dtype=torch.int16, device=dev)
    d_end = torch.empty((n, 2), dtype=torch.int16, device=dev)

    def step_resident():
        ctx.align_device(NW, POLICY_DEFAULT_OCL, d_reads, d_refs, d_a, d_b, d_start, d_end, SCORING, stream=stream)

    # clocks: nvidia-smi is started before the warm-up (it needs a second to come up, longer on an 8-GPU box) and samples
    # every 50 ms through the warm-up, the timed steps and the profiled pass of the same steps
    sampler = ClockSampler(dev_index)
    if RANK == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    launches_per_step = ctx.timings()["launches"]
    if RANK == 0:
        t_wait = time.time()
        while not sampler.lines and time.time() - t_wait < 5.0:  # first sample in before the timed region starts
            step_resident()
            torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    resident_ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    # per-kernel times for the roofline: the same K steps once more with CUDA events around the prep, fill and
    # traceback kernels on the same stream (read back step by step, so kept out of the timed region above)
    ctx.set_profiling(True)
    kernel_ms = [0.0, 0.0, 0.0]
    for _ in range(args.steps):
        step_resident()
        k = ctx.kernel_ms()
        kernel_ms = [a + b for a, b in zip(kernel_ms, k)]
    ctx.set_profiling(False)
    barrier()
    clocks = sampler.stop() if RANK == 0 else None
    value = WORLD * cells_per_step / (resident_ms * 1e-3) / 1e9
    fill_ms = kernel_ms[1] / args.steps
    prep_ms = kernel_ms[0] / args.steps
    tb_ms = kernel_ms[2] / args.steps

    # spot-check the resident result against the plug-in path later (same inputs)
    start_resident = d_start.cpu().numpy().copy()

    # ---------------- every function on the same shape (resident), incl. the SSE/AVX pointer policy ----------------
    n_modes = min(n, 250_000)
    mr, mf = d_reads[:n_modes], d_refs[:n_modes]
    m_s = torch.zeros(n_modes, dtype=torch.int16, device=dev)
    mcells = n_modes * CELLS_PER_PAIR
    modes = {}
    for name, fn in (
            ("sw_score", lambda: ctx.score_device(SW, mr, mf, m_s, SCORING, stream=stream)),
            ("nw_score", lambda: ctx.score_device(NW, mr, mf, m_s, SCORING, stream=stream)),
            ("sw_align_default_ocl", lambda: ctx.align_device(SW, POLICY_DEFAULT_OCL, mr, mf, d_a, d_b, d_start, d_end, SCORING, stream=stream)),
            ("nw_align_default_ocl", lambda: ctx.align_device(NW, POLICY_DEFAULT_OCL, mr, mf, d_a, d_b, d_start, d_end, SCORING, stream=stream)),
            ("sw_align_simd", lambda: ctx.align_device(SW, POLICY_SIMD, mr, mf, d_a, d_b, d_start, d_end, SCORING, stream=stream)),
            ("nw_align_simd", lambda: ctx.align_device(NW, POLICY_SIMD, mr, mf, d_a, d_b, d_start, d_end, SCORING, stream=stream)),
            # the 32-bit general kernel (what leaves the packed kernels' domain falls to it), here through the affine-gap
            # variant, which has no packed kernel yet
            ("sw_score_affine_general_kernel", lambda: ctx.score_device(capi.affine_opt(SW, -5), mr, mf, m_s, SCORING, stream=stream)),
            ("nw_align_affine_general_kernel", lambda: ctx.align_device(capi.affine_opt(NW, -5), POLICY_DEFAULT_OCL, mr, mf, d_a, d_b, d_start, d_end, SCORING, stream=stream))):
        ms = timed_resident(torch, fn, 3)
        modes[name] = {"resident_gcups": mcells / ms / 1e6, "ms": ms}
    modes["_note"] = f"{n_modes} pairs of the C2 shape, prep + fill (+ traceback), this rank"

    # ---------------- end to end through the plug-in boundary, host buffers ----------------
    host_threads = max(1, (os.cpu_count() or 1) // max(1, WORLD))
    e2e_times = []
    h = PluginHost(capi.library_path(), READ_LEN, REF_LEN, SCORING, num_threads=host_threads,
                   extra={"cuda_devices": 1, "cuda_device_first": dev_index}, verbosity=0)
    h.stage(reads, refs, scattered=True)
    for it in range(args.warmup + args.steps):
        barrier()
        h.align_staged(NW, fetch=False)
        # steady_clock around the virtual compute_alignments call itself (csrc/plugin_host.cpp); the
        # caller-side Alignment[n] array is allocated before it, like main.cpp:123 does
        if it >= args.warmup:
            e2e_times.append(max_over_ranks(h.last_call_seconds))
        if it == args.warmup + args.steps - 1:
            _, _, fields = h.fetch_alignments()
            if not np.array_equal(fields[:, 0], start_resident):
                raise SystemExit("bench.py: plug-in path and resident path disagree")
        h.drop_alignments()
    pt = capi.plugin_timings() or {}
    h.close()
    e2e_sec = sum(e2e_times) / len(e2e_times)
    e2e_value = WORLD * cells_per_step / e2e_sec / 1e9

    # ---------------- the batch-friendly entry point beside it ----------------
    # same kernels, same alignments; offset-addressed sequences in page-locked host memory in, CIGARs + coordinates out
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    pk_reads, pk_ro = synth.pack_batch(reads)
    pk_refs, pk_fo = synth.pack_batch(refs)
    pk_reads, pk_ro, pk_refs, pk_fo = pin(pk_reads), pin(pk_ro), pin(pk_refs), pin(pk_fo)
    ctx.set_host_threads(host_threads)
    pk_times = []
    pk_out: dict = {}  # output arrays are the caller's and are reused across steps, like the Alignment[n] array of the legacy call
    for it in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        pk_scores, pk_coords, pk_coff, pk_cigar = ctx.align_packed(NW, POLICY_DEFAULT_OCL, pk_reads, pk_ro, pk_refs, pk_fo, SCORING, out=pk_out)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            pk_times.append(max_over_ranks(dt))
    pk_t = ctx.timings()
    # same paths as the plug-in call: aligned columns = L - 1 - start
    if not np.array_equal((pk_cigar >> 4).astype(np.int64).sum(), np.int64((L - 1 - start_resident.astype(np.int64)).sum())):
        raise SystemExit("bench.py: packed path and resident path disagree")
    pk_sec = sum(pk_times) / len(pk_times)
    e2e_packed = {"value": WORLD * cells_per_step / pk_sec / 1e9, "unit": "GCUPS", "ms_per_step": pk_sec * 1e3,
                  "h2d_bytes_per_step": int(pk_t["h2d_bytes"]), "d2h_bytes_per_step": int(pk_t["d2h_bytes"]),
                  "cigar_ops": int(pk_coff[-1]), "phases": pk_t, "limiter": limiter_of(pk_t, pk_sec * 1e3),
                  "fraction_of_resident": (resident_ms / (pk_sec * 1e3)),
                  "api": "va_cuda_align_packed: contiguous reads/refs + offsets in page-locked host memory in, scores + coordinates + BAM-style CIGARs out (host)"}
    del pk_reads, pk_refs, pk_out

    # ---------------- roofline of the dominant (fill) kernel ----------------
    peaks = measured_peaks()
    int_peak = ctx.int_peak(1, stream=stream)      # VIADDMNMX.S16x2 lane-ops/s, measured now
    torch.cuda.synchronize()
    roof_gcups = int_peak / 3.0 / 1e9
    fill_gcups = cells_per_step / (fill_ms * 1e-3) / 1e9 if fill_ms > 0 else 0.0
    # algorithmic HBM bytes per pair of the fill launch: codes in (read+ref bytes), 2-bit directions
    # out (rows x cols / 4), end cell + score out (6 B)
    fill_bytes = n * (READ_LEN + REF_LEN + CELLS_PER_PAIR / 4 + 6)
    hbm_achieved = fill_bytes / (fill_ms * 1e-3) / 1e9 if fill_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("fill_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "int_alu", "kernel": "fill (DP recurrence + 2-bit directions)", "achieved": fill_gcups,
        "peak": roof_gcups, "unit": "GCUPS", "frac": fill_gcups / roof_gcups if roof_gcups else None,
        "peak_source": "VIADDMNMX.S16x2 lane-ops/s measured in this run (va_cuda_int_peak) / 3 ops per cell",
        "ms_per_launch": {"prep": prep_ms, "fill": fill_ms, "traceback": tb_ms},
        "whole_step_frac": (cells_per_step / (resident_ms * 1e-3) / 1e9) / roof_gcups if roof_gcups else None,
        "traffic": traffic,
        "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": hbm_achieved / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                "algorithmic_bytes_per_launch": fill_bytes, "peak_source": peaks["_source"]},
    }
    for m in modes.values():
        if isinstance(m, dict):
            m["roofline_frac"] = m["resident_gcups"] / roof_gcups

    # ---------------- the other BASELINE configs ----------------
    del d_a, d_b, d_reads, d_refs, mr, mf
    torch.cuda.empty_cache()
    configs = {}
    if not args.no_configs:
        barrier()
        if RANK == 0:
            try:
                configs["C1"] = config_c1(torch, ctx, dev, roof_gcups)
            except Exception as ex:  # evidence section: a failure here must not lose the headline line
                configs["C1"] = {"error": repr(ex)}
        barrier()
        for tag, fn in (("C3", config_c3), ("C4", config_c4)):
            try:
                configs[tag] = fn(torch, ctx, dev, roof_gcups, max_over_ranks, barrier, args.quick)
            except Exception as ex:
                configs[tag] = {"error": repr(ex)}
                barrier()
            torch.cuda.empty_cache()

    # ---------------- (N > 1) the in-process shard: ONE compute_alignments call over all N devices on rank 0 ----------------
    e2e_inprocess = None
    ctx.close()
    if distributed and not args.no_inprocess:
        torch.cuda.empty_cache()
        barrier()
        if RANK == 0:
            try:
                big_r, big_f = np.tile(reads, (WORLD, 1)), np.tile(refs, (WORLD, 1))
                hp = PluginHost(capi.library_path(), READ_LEN, REF_LEN, SCORING, num_threads=os.cpu_count() or 1,
                                extra={"cuda_devices": WORLD, "cuda_device_first": 0}, verbosity=0)
                hp.stage(big_r, big_f, scattered=True)
                times = []
                for it in range(2 + min(args.steps, 5)):
                    hp.align_staged(NW, fetch=False)
                    if it >= 2:
                        times.append(hp.last_call_seconds)
                    hp.drop_alignments()
                ipt = capi.plugin_timings() or {}
                hp.close()
                sec = sum(times) / len(times)
                e2e_inprocess = {"value": WORLD * cells_per_step / sec / 1e9, "unit": "GCUPS", "ms_per_step": sec * 1e3,
                                 "devices": WORLD, "host_threads": os.cpu_count() or 1, "phases": ipt, "limiter": limiter_of(ipt, sec * 1e3),
                                 "api": "ONE compute_alignments call on rank 0, cuda_devices=N: pair range sharded over the N GPUs inside the plug-in"}
            except Exception as ex:
                e2e_inprocess = {"error": repr(ex)}
        barrier()

    cpu_baseline = None
    if RANK == 0 and WORLD == 1 and not args.no_cpu:
        cb = reference_arm(1, 1, REFERENCE_SAMPLE_PAIRS, os.cpu_count() or 1)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if RANK == 0:
        line = {
            "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": WORLD, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": resident_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": config_dict(n, WORLD),
            "e2e": {"value": e2e_value, "unit": "GCUPS", "ms_per_step": e2e_sec * 1e3,
                    "h2d_bytes_per_step": n * (READ_LEN + REF_LEN), "d2h_bytes_per_step": int(pt.get("d2h_bytes", n * (2 * L + 6))),
                    "host_threads": host_threads, "host_malloc": "plug-in: mallopt keeps freed pages (VERSALIGN_CUDA_MALLOC_TUNE)" if not args.tune_malloc else args.host_malloc,
                    "phases": pt, "limiter": limiter_of(pt, e2e_sec * 1e3),
                    "api": "dlopen(libCUDAKernel.so) -> spawn_alignment_kernel -> AlignmentKernel::compute_alignments, scattered char* in, new char[] out"},
            "e2e_packed": e2e_packed,
            "e2e_inprocess": e2e_inprocess,
            "gpu_launches": int(launches_per_step) * args.steps,
            "roofline": roofline, "modes": modes, "configs": configs, "cpu_baseline": cpu_baseline, "clocks": clocks,
        }
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU (default: the C2 workload, 1M)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C3/C4 evidence section")
    ap.add_argument("--no-inprocess", action="store_true", help="skip the in-process N-device call on rank 0")
    ap.add_argument("--quick", action="store_true", help="configs at 1/10 size (development)")
    ap.add_argument("--tune-malloc", action="store_true", help="host side: keep freed pages in glibc's arenas (mallopt) in this process too")
    args = ap.parse_args()
    args.host_malloc = tune_host_malloc() if args.tune_malloc else "default"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
