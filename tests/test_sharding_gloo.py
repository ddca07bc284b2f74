"""The N>1 path on CPU: two gloo ranks shard a batch the way bench.py / the C ABI do, align their
slices (with the oracle standing in for the device -- there is no GPU here), gather to rank 0 and
must reproduce the single-process result; timings reduce with max-over-ranks."""
import os
import socket
import subprocess
import sys

import numpy as np

from versalignlib_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["VA_ROOT"])
from oracle import binding as ora
from versalignlib_b200 import shard, synth

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 301
reads, refs, rl, fl = synth.mixed_batch(n, 20, 60, p_sub=0.1, q_indel=0.02, seed=5)
bounds = shard.cell_balanced_bounds(rl, fl, world)
lo, hi = bounds[rank]
local = ora.score(ora.SW, np.ascontiguousarray(reads[lo:hi]), np.ascontiguousarray(refs[lo:hi]), threads=1)
full = shard.gather_slices(local, n, bounds)
a, b, st, _ = ora.align(ora.NW, 0, np.ascontiguousarray(reads[lo:hi]), np.ascontiguousarray(refs[lo:hi]), threads=1)
full_st = shard.gather_slices(st, n, bounds)
full_a = shard.gather_slices(a, n, bounds)
t = shard.max_over_ranks(1.0 + rank)
assert t == float(world), t
dist.barrier()
if rank == 0:
    want = ora.score(ora.SW, reads, refs, threads=1)
    wa, wb, wst, _ = ora.align(ora.NW, 0, reads, refs, threads=1)
    assert np.array_equal(full, want)
    assert np.array_equal(full_st, wst) and np.array_equal(full_a, wa)
    print("gloo shard ok", bounds)
dist.destroy_process_group()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_bounds_cover_and_balance():
    for n in (0, 1, 7, 100, 1001):
        for w in (1, 2, 3, 8):
            b = shard.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    rng = np.random.default_rng(1)
    rows, cols = rng.integers(100, 251, 5000), rng.integers(100, 251, 5000)
    b = shard.cell_balanced_bounds(rows, cols, 8)
    cells = [int((rows[lo:hi].astype(np.int64) * cols[lo:hi]).sum()) for lo, hi in b]
    assert b[0][0] == 0 and b[-1][1] == 5000
    assert max(cells) / (sum(cells) / 8) < 1.02


def test_two_gloo_ranks_reproduce_single_process_result():
    port = _free_port()
    env = dict(os.environ, VA_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-c", WORKER]
    # torchrun has no -c: write the worker to a temp file instead
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix="_va_gloo_worker.py", delete=False) as f:
        f.write(WORKER)
        path = f.name
    try:
        cmd = cmd[:-2] + [path]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert r.returncode == 0 and "gloo shard ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    finally:
        os.unlink(path)
