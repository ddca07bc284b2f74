"""Sharding of a pair batch over ranks / devices.  Pairs are independent (the reference's only
parallel axis, DefaultKernel.cpp:45-48), so there is no data-path collective: every rank aligns a
contiguous slice and results land in disjoint slices of the caller's arrays.  The only
communication is control-plane: a barrier and a max-reduction of the timings, and (optionally) a
gather of the results to rank 0.

The same contiguous split is what the C ABI uses inside one process for the devices of a context
(csrc/va_cabi.cu run_host_call)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Equal pair counts: slice r is [n*r//world, n*(r+1)//world)."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def cell_balanced_bounds(rows: np.ndarray, cols: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous slices with (almost) equal DP cells, for mixed-length batches."""
    cells = rows.astype(np.int64) * cols.astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(cells)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(csum, total * r / world, side="left")))
    cuts.append(len(cells))
    cuts = np.maximum.accumulate(np.clip(cuts, 0, len(cells)))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_slices(local: np.ndarray, n_total: int, bounds: list[tuple[int, int]]):
    """Gather per-rank result slices to rank 0 (returns the full array there, None elsewhere)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    rank, world = dist.get_rank(), dist.get_world_size()
    longest = max(hi - lo for lo, hi in bounds)
    pad = np.zeros((longest,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad.view(np.uint8).reshape(-1))  # bytes: gloo has no 16-bit integer types
    out = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, out, dst=0)
    if rank != 0:
        return None
    full = np.zeros((n_total,) + local.shape[1:], dtype=local.dtype)
    for r, (lo, hi) in enumerate(bounds):
        full[lo:hi] = out[r].numpy().view(local.dtype).reshape(pad.shape)[: hi - lo]
    return full
