#!/usr/bin/env python
"""Throughput of the "next" kernels on the config-2 population (synthetic, 1000
structures x 29 838 beads): K3 restraint selection over the records of an A-step,
K4 / K4b SPRITE Rg^2 over random clusters, K5 rank matching of every polymer bond, DamID activation distances of every locus,
haploid contact map.  Kernel-only times (CUDA events inside the library)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nstruct", type=int, default=1000)
    ap.add_argument("--pairs", type=int, default=400000)
    args = ap.parse_args()
    import torch
    from igm_b200 import synthetic
    from igm_b200.engine import ActdistEngine
    from igm_b200.steps.ActivationDistanceStep import filter_candidates
    dev = torch.device("cuda:0")
    bins = synthetic.genome_bins(200_000)
    chrom_hap, chrom_bead, copy_bead, ci = synthetic.build_index(bins)
    nbead = len(chrom_bead)
    radius = float(synthetic.bead_radius(nbead))
    coords = synthetic.random_walk_coordinates_torch(chrom_bead, copy_bead, args.nstruct, radius, 20261018, dev)
    eng = ActdistEngine(nbead=nbead, nstruct=args.nstruct, device=0)
    eng.upload_coordinates(coords)
    eng.set_index(ci.ptr, ci.beads, chrom_hap, np.full(nbead, radius, np.float32))
    eng.set_bead_chrom(chrom_bead)
    out = {}
    # A-step on a slice of the candidate list -> records for K3
    pm = synthetic.make_prob_matrix(chrom_hap, seed=20261018)
    ii, jj, pw = filter_candidates(pm, 0.01, 0.01)
    ii, jj, pw = ii[:args.pairs], jj[:args.pairs], pw[:args.pairs]
    res = eng.actdist(ii, jj, pw)
    row, col, dist, prob = eng.expand_records(ii, jj, res)
    for kind in ("intra", "inter"):
        bitmap, counts = eng.restraint_select(row, col, dist, kind)
        ms = eng.last_kernel_ms()
        out["K3_" + kind] = {"records": int(len(row)), "ms": ms,
                             "record_structs_per_s": len(row) * args.nstruct / (ms * 1e-3),
                             "GBps_algorithmic": len(row) * (24.0 * args.nstruct) / (ms * 1e-3) / 1e9,
                             "assigned_fraction": float(counts.sum()) / (len(row) * args.nstruct)}
    # K4: 20 000 random clusters of 2..6 loci
    rng = np.random.default_rng(1)
    clusters = []
    for _ in range(20000):
        loci = rng.choice(len(chrom_hap), size=int(rng.integers(2, 7)), replace=False)
        clusters.append([ci[int(l)] for l in loci])
    eng.sprite_rg2(clusters)
    ms = eng.last_kernel_ms()
    out["K4"] = {"clusters": len(clusters), "ms": ms, "cluster_structs_per_s": len(clusters) * args.nstruct / (ms * 1e-3)}
    # K4b: whole-cluster Rg^2 under a per-structure copy choice: 2 000 clusters of 20..200 segments
    full, nseg_tot = [], 0
    for _ in range(2000):
        loci = rng.choice(len(chrom_hap), size=int(rng.integers(20, 201)), replace=False)
        groups = (np.arange(len(loci)) % 4).tolist()
        sel = rng.integers(0, 2, (args.nstruct, 4)).astype(np.int32)
        # haploid loci have one copy: selection 1 would be out of range -> use the last copy (-1)
        sel[sel == 1] = -1
        full.append(([ci[int(l)] for l in loci], groups, sel))
        nseg_tot += len(loci)
    eng.sprite_cluster_rg2(full)
    ms = eng.last_kernel_ms()
    out["K4b"] = {"clusters": len(full), "segments": nseg_tot, "ms": ms,
                  "segment_structs_per_s": nseg_tot * args.nstruct / (ms * 1e-3),
                  "GBps_algorithmic": nseg_tot * 12.0 * args.nstruct / (ms * 1e-3) / 1e9}
    # DamID: every locus
    nh = len(chrom_hap)
    pe = rng.uniform(0.05, 1, nh).astype(np.float32)
    eng.damid_actdist(np.arange(nh), pe, None, 5000.0, 0.05, 1)
    ms = eng.last_kernel_ms()
    out["DamID"] = {"loci": nh, "ms": ms, "loci_per_s": nh / (ms * 1e-3)}
    # K5: every polymer bond of the population with per-bond targets (PolymerAssignmentStep)
    bonds = np.arange(nbead - 1, dtype=np.int32)
    neg = np.full(nbead - 1, -1, np.int32)
    tgt = np.sort(rng.uniform(50, 900, args.nstruct)).astype(np.float32)
    eng.rank_match(np.stack([bonds, neg], 1), np.stack([bonds + 1, neg], 1), "min", tgt, want_rank=False, want_value=False)
    ms = eng.last_kernel_ms()
    out["K5_polymer"] = {"bonds": int(nbead - 1), "ms": ms, "bond_structs_per_s": (nbead - 1) * args.nstruct / (ms * 1e-3)}
    # haploid contact map, first 2048 rows against all columns
    eng.contact_counts_haploid(0, 2048, 0, nh)
    ms = eng.last_kernel_ms()
    out["K2_haploid"] = {"locus_pairs": 2048 * nh, "ms": ms,
                         "bead_pair_structs_per_s": 2048 * nh * 4.0 * args.nstruct / (ms * 1e-3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
