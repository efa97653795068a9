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

import os
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


def bead_shares(nbead: int, world: int) -> Tuple[int, List[Tuple[int, int]]]:
    """Equal contiguous bead ranges, one per GPU (the last may be short or empty)."""
    return shard_bounds(nbead, world)


def replicate_population(eng, xyz_host, rank: int, world: int, group=None) -> None:
    """"Coordinates replicated per GPU" without one PCIe upload per GPU: every rank uploads
    only ITS share of the beads from host memory (1 / world of the host-side traffic), then
    one all-gather of whole staged rows over NVLink / NVSwitch completes every GPU's copy
    (in place in the engine's own buffer).  ``xyz_host``: (nbead, nstruct, 3) float32, the
    .hss layout, ideally page-locked; every rank passes the same population."""
    import torch.distributed as dist
    per, shares = bead_shares(eng.nbead, world)
    lo, hi = shares[rank]
    if hi > lo:
        eng.upload_coordinates(xyz_host[lo:hi], bead0=lo)
    if world == 1:
        return
    buf = eng.coords_tensor()                           # (nbead + 1 + spare rows, 3 * npad)
    if world * per > buf.shape[0]:
        raise ValueError("too many ranks for the staged buffer's spare rows")
    whole = buf[:world * per]
    # rows past nbead are zero on every rank (never uploaded), so the tail of the last share
    # gathers zeros over zeros: the all-zero origin row stays intact
    mine = whole[rank * per:(rank + 1) * per]
    if dist.get_backend(group) != "nccl":
        mine = mine.clone()
    dist.all_gather_into_tensor(whole.view(-1), mine.reshape(-1), group=group)


class PeerGather:
    """Gather buffers the A-step kernel writes into directly over NVLink.

    Every rank owns `nbuf` buffers of shape (world, per, 32) uint8 in peer-mapped
    ("symmetric") device memory.  igmk_actdist_device_peers stores each raw pair
    result of rank r into slice [r] of the current buffer of EVERY rank from inside
    the kernel - finished, dist / prob included - so by the time the kernel ends the
    all-gather has already happened; what is left is one cross-rank barrier.
    Alternating between two buffers makes that single barrier per step sufficient:
    a rank can only start writing buffer b again after every rank has passed the
    barrier of the step that last read it.

    torch supplies the plumbing (symmetric allocation, rendezvous, barrier)."""

    def __init__(self, per: int, rank: int, world: int, device, group=None, nbuf: int = 2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        grp = group if group is not None else dist.group.WORLD
        name = grp.group_name
        if hasattr(symm, "enable_symm_mem_for_group"):
            try:
                symm.enable_symm_mem_for_group(name)
            except Exception:
                pass
        self.rank, self.world, self.per = rank, world, max(per, 1)
        self.bufs, self.hdls, self.peer_slices = [], [], []
        for _ in range(nbuf):
            buf = symm.empty((world, self.per, RESULT_BYTES), dtype=torch.uint8, device=device)
            buf.zero_()
            hdl = symm.rendezvous(buf, name)
            off = rank * self.per * RESULT_BYTES
            self.bufs.append(buf)
            self.hdls.append(hdl)
            self.peer_slices.append(torch.tensor([int(p) + off for p in hdl.buffer_ptrs],
                                                 dtype=torch.int64, device=device))
        self.k = 0
        torch.cuda.synchronize(device)
        dist.barrier(group=grp)

    def step(self, eng, d_i, d_j, d_pw, d_pl, n_pairs: int, contact_range=2.0, it_corr=0, mode="LB",
             stream: int = 0):
        """One sharded A-step: kernel with peer stores -> barrier.  Returns
        the (world, per, 32) uint8 tensor holding every rank's results (valid once the
        stream has been synchronised)."""
        b = self.k % len(self.bufs)
        self.k += 1
        eng.actdist_device_peers(d_i, d_j, d_pw, d_pl, self.peer_slices[b], self.world, n_pairs,
                                 contact_range, it_corr, mode, stream)
        self.hdls[b].barrier()           # the records arrive finished (dist / prob filled in by the kernel)
        return self.bufs[b]


def peer_gather_available(world: int = 0) -> bool:
    """In-kernel peer stores beat NCCL's all-gather from 4 GPUs up (measured on
    8 x B200, config 2: 1.648 vs 1.608 G pairs/s; at 2 GPUs 431 vs 440 M)."""
    if world and world < 4 and os.environ.get("IGMK_PEER_GATHER", "") != "1":
        return False
    try:
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory  # noqa: F401
        return (torch.cuda.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
                and os.environ.get("IGMK_NO_PEER_GATHER", "0") != "1")
    except Exception:
        return False


def gathered_to_results(full, per: int, bounds) -> np.ndarray:
    """Host structured array of all pairs in input order (padding rows dropped)."""
    host = full.cpu().numpy()
    parts = [host[r, :hi - lo] for r, (lo, hi) in enumerate(bounds)]
    flat = np.ascontiguousarray(np.concatenate(parts, axis=0)) if parts else np.zeros((0, RESULT_BYTES), np.uint8)
    return flat.reshape(-1).view(PAIR_RESULT_DTYPE)
