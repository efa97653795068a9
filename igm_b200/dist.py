"""Multi-GPU sharding of the A-step: one process per GPU (torch.distributed,
NCCL over NVLink/NVSwitch), coordinates replicated, candidate pairs split into
contiguous equal-count shards, one all-gather of the fixed-width per-pair
results.  The reference farms 1000-pair batches to CPU workers
through ipyparallel and the shared filesystem (igm/parallel/,
igm/steps/ActivationDistanceStep.py:181-194, igm/core/step.py:274); pairs are
independent there too, so there is no other exchange step.

torch supplies process-group plumbing and device buffers only; the kernel
writes each rank's results straight into its slice of the all-gather buffer.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np

from ._lib import PAIR_RESULT_DTYPE

RESULT_BYTES = PAIR_RESULT_DTYPE.itemsize   # 32


def shard_bounds(n_pairs: int, world: int) -> Tuple[int, List[Tuple[int, int]]]:
    """Contiguous shards of `per` = ceil(n / world) pairs (the last may be short
    or empty).  Contiguity keeps locus-i locality (CSR order) and lets the
    gathered buffer be read back in the reference's record order."""
    per = (n_pairs + world - 1) // world if world > 0 else 0
    out = []
    for r in range(world):
        lo = min(n_pairs, r * per)
        out.append((lo, min(n_pairs, lo + per)))
    return per, out


def run_sharded(compute_shard: Callable, n_pairs: int, rank: int, world: int, device,
                group=None):
    """compute_shard(lo, hi, out_tensor) must fill out_tensor ((hi-lo), 32) uint8 on
    `device` with igmk_pair_result records for pairs [lo, hi) (asynchronously on
    the current stream is fine).  Returns a (world * per, 32) uint8 tensor holding
    every rank's results, shard r at rows [r * per, r * per + count_r)."""
    import torch
    import torch.distributed as dist
    per, bounds = shard_bounds(n_pairs, world)
    full = torch.zeros((world, max(per, 1), RESULT_BYTES), dtype=torch.uint8, device=device)
    lo, hi = bounds[rank]
    if hi > lo:
        compute_shard(lo, hi, full[rank, :hi - lo])
    if world > 1:
        flat = full.view(world * max(per, 1), RESULT_BYTES)
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(flat, full[rank], group=group)      # in place
        else:
            dist.all_gather_into_tensor(flat, full[rank].clone(), group=group)
    return full, per, bounds


def gathered_to_results(full, per: int, bounds) -> np.ndarray:
    """Host structured array of all pairs in input order (padding rows dropped)."""
    host = full.cpu().numpy()
    parts = [host[r, :hi - lo] for r, (lo, hi) in enumerate(bounds)]
    flat = np.ascontiguousarray(np.concatenate(parts, axis=0)) if parts else np.zeros((0, RESULT_BYTES), np.uint8)
    return flat.reshape(-1).view(PAIR_RESULT_DTYPE)
