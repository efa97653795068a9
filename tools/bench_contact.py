#!/usr/bin/env python
"""Config 4 of BASELINE.json: population contact-frequency counts (K2) for a
synthetic 1000-structure x 29 838-bead population, all bead pairs, computed tile
row by tile row (upper triangle at block granularity).  Reports bead-pair-structs/s
and the fraction of the FP32 lane roofline (9 lane-operations per bead pair per
structure: 8 non-FMA float32 operations + 1 compare; peak = SMs x 128 lanes x SM
clock - quoted at the maximum clock and at the median clock sampled during the
run)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nstruct", type=int, default=1000)
    ap.add_argument("--resolution", type=int, default=200_000)
    ap.add_argument("--block", type=int, default=4096)
    ap.add_argument("--max-blocks", type=int, default=0, help="only the first K block rows (debug)")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from igm_b200 import synthetic
    from igm_b200.contact import row_blocks
    from igm_b200.engine import ActdistEngine, launch_count
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bins = synthetic.genome_bins(args.resolution)
    chrom_hap, chrom_bead, copy_bead, ci = synthetic.build_index(bins)
    nbead = len(chrom_bead)
    radius = float(synthetic.bead_radius(nbead))
    coords = synthetic.random_walk_coordinates_torch(chrom_bead, copy_bead, args.nstruct, radius, 7, dev)
    eng = ActdistEngine(nbead=nbead, nstruct=args.nstruct, device=lr)
    eng.upload_coordinates(coords)
    eng.set_index(ci.ptr, ci.beads, chrom_hap, np.full(nbead, radius, np.float32))
    B = args.block
    out = torch.zeros((B, nbead), dtype=torch.int32, device=dev)
    nblk = (nbead + B - 1) // B
    if args.max_blocks:
        nblk = min(nblk, args.max_blocks)
    stream = torch.cuda.current_stream().cuda_stream
    # warm-up
    eng.contact_counts_device(0, min(B, nbead), 0, min(B, nbead), out, 2.0, False, stream)
    torch.cuda.synchronize()
    from bench import ClockSampler
    # block rows of the upper triangle dealt to the ranks (no collective: tiles are independent)
    mine = row_blocks(nbead, B, rank, world)
    if args.max_blocks:
        mine = mine[:args.max_blocks]
    sampler = ClockSampler(lr)
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = launch_count()
    sampler.mark_start()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.reps):
        pairs = 0
        for r0, r1 in mine:
            nr = r1 - r0
            # columns from the start of this block row to the end (upper triangle, block granularity)
            eng.contact_counts_device(r0, nr, r0, nbead - r0, out, 2.0, False, stream)
            pairs += nr * (nbead - r0)
    ev1.record()
    torch.cuda.synchronize()
    sampler.mark_stop()
    time.sleep(0.1)
    sampler.stop()
    clocks = sampler.summary()
    ms = ev0.elapsed_time(ev1) / args.reps
    if world > 1:                                   # whole job: all pairs / slowest rank
        t = torch.tensor([ms, float(pairs)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, pairs = float(tmax[0].item()), int(t[1].item())
    ops = pairs * args.nstruct
    sm = sm_count = torch.cuda.get_device_properties(lr).multi_processor_count * world
    peak = sm * 128 * 1.965e9          # FP32 lanes x max SM clock: non-FMA instr/s
    res = {"metric": "contact-frequency bead-pair-structs/s", "value": ops / (ms * 1e-3), "ms": ms,
           "bead_pairs": pairs, "nstruct": args.nstruct, "nbead": nbead,
           "roofline": {"bound": "fp32-lanes", "achieved_laneops_per_s": 9 * ops / (ms * 1e-3),
                        "peak_laneops_per_s": peak, "frac": 9 * ops / (ms * 1e-3) / peak},
           "n_gpus": world, "gpu_launches": launch_count() - l0, "reps": args.reps, "clocks": clocks}
    if clocks.get("sm_mhz"):
        res["roofline"]["frac_at_sampled_clock"] = 9 * ops / (ms * 1e-3) / (sm * 128 * clocks["sm_mhz"] * 1e6)
    if args.check:
        from oracle import contact_oracle as co
        h = coords[:64].cpu().numpy()
        got = torch.zeros((64, 64), dtype=torch.int32, device=dev)
        eng.contact_counts_device(0, 64, 0, 64, got, 2.0, False, stream)
        torch.cuda.synchronize()
        exp = co.contact_counts_fast(h, np.full(64, radius, np.float32), np.arange(64), np.arange(64))
        res["parity_sample_ok"] = bool(np.array_equal(got.cpu().numpy().astype(np.uint32), exp))
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
