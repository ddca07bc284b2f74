#!/usr/bin/env python
"""bench.py -- GCUPS of the batched DP hot path on B200, one process per GPU.

Workload (BASELINE.json configs[1], "C2"): the library's "Needleman-Wunsch" compute_alignments
(score + traceback, gapped strings out) on 1 M synthetic 150 bp read/ref pairs PER GPU
(weak scaling: pairs are independent, no collective on the data path).

  value      whole-job GCUPS with the raw batch already resident in HBM: prep + fill +
             traceback kernels through va_cuda_align_device on torch's current stream,
             CUDA events, max over ranks
  e2e        the same metric through the reference-facing plug-in boundary
             (dlopen -> spawn_alignment_kernel -> AlignmentKernel::compute_alignments) with
             scattered HOST buffers in and new char[] blocks out; H2D/D2H inside the timed region
  roofline   the fill kernel against the measured integer-pipe peak (VIADDMNMX.S16x2 lane-ops/s
             divided by 3 ops per cell), plus its HBM view
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
NW = 1
POLICY_DEFAULT_OCL = 0
CELLS_PER_PAIR = READ_LEN * REF_LEN
WORKLOAD = "C2: NW compute_alignments (score+traceback), 1M x (150bp read vs 150bp ref) per GPU, p_sub=0.08 q_indel=0.02"


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
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
    """Host-application side: keep freed heap pages inside glibc's arenas between calls.  Every
    kernel behind this boundary (the reference's too) returns 2 heap blocks per pair, and the
    caller frees them after each call; with the default trim threshold each call then re-faults
    hundreds of MB.  Applied to BOTH arms (ours and --impl reference)."""
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
    # ~0.3-1 GCUPS expected: 48k pairs x 22.5k cells ~ 1.1 Gcells -> a few seconds per step
    cb = reference_arm(args.steps, args.warmup, 48_000, threads)
    line = {
        "impl": "reference", "metric": "GCUPS", "value": cb["value"], "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU reference arm: bounded sample, rank 0 only",
                   "host_malloc": args.host_malloc},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from versalignlib_b200 import capi
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

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    launches_per_step = ctx.timings()["launches"]
    sampler = ClockSampler(dev_index)
    if RANK == 0:
        sampler.start()
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
    # same kernels, same alignments; offset-addressed sequences in, CIGARs + coordinates out
    from versalignlib_b200 import synth as _synth
    pk_reads, pk_ro = _synth.pack_batch(reads)
    pk_refs, pk_fo = _synth.pack_batch(refs)
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
                  "cigar_ops": int(pk_coff[-1]), "phases": pk_t,
                  "api": "va_cuda_align_packed: contiguous reads/refs + offsets in (host), scores + coordinates + BAM-style CIGARs out (host)"}

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
        "traffic": traffic,
        "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": hbm_achieved / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                "algorithmic_bytes_per_launch": fill_bytes, "peak_source": peaks["_source"]},
    }

    cpu_baseline = None
    if RANK == 0 and WORLD == 1 and not args.no_cpu:
        cb = reference_arm(1, 1, 48_000, os.cpu_count() or 1)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if RANK == 0:
        line = {
            "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": WORLD, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": resident_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": n, "read_length": READ_LEN, "ref_length": REF_LEN,
                       "scoring": list(SCORING), "traceback_policy": "DEFAULT_OCL", "parallelism": f"dp{WORLD} (independent pairs, no collective)",
                       "l2": "inputs + direction matrix (5.9 GB/step) exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "GCUPS", "ms_per_step": e2e_sec * 1e3,
                    "h2d_bytes_per_step": n * (READ_LEN + REF_LEN), "d2h_bytes_per_step": int(pt.get("d2h_bytes", n * (2 * L + 6))),
                    "host_threads": host_threads, "host_malloc": args.host_malloc, "phases": pt,
                    "api": "dlopen(libCUDAKernel.so) -> spawn_alignment_kernel -> AlignmentKernel::compute_alignments, scattered char* in, new char[] out"},
            "e2e_packed": e2e_packed,
            "gpu_launches": int(launches_per_step) * args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
        }
        print(json.dumps(line))
    ctx.close()
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
    ap.add_argument("--tune-malloc", action="store_true", help="host side: keep freed pages in glibc's arenas (mallopt)")
    args = ap.parse_args()
    args.host_malloc = tune_host_malloc() if args.tune_malloc else "default"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
