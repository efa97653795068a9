#!/usr/bin/env python
"""Whole Hi-C A-step through the Step drop-in (setup -> task -> reduce) on config 2 of
BASELINE.json, files included: a synthetic .hss (1000 structures x 29 838 beads) and .hcs
are written to a scratch directory, then igm_b200.steps.ActivationDistanceStep(cfg).run()
is timed phase by phase for every sigma of the demo sweep (1.0 -> 0.01, iterative
correction on, each iteration reading the previous actdist.hdf5).  What `igm-run` would see (bin/igm-run:164-166)."""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


# the sigma sweep of demo/config_file.json:35-50 (config 2: "sigma sweep 1.0 -> 0.01")
SIGMAS = (1.0, 0.2, 0.1, 0.05, 0.02, 0.01)


def main():
    import torch
    from igm_b200 import hdf5, synthetic
    from igm_b200.population import CopyIndex, Population
    from igm_b200.steps import ActivationDistanceStep
    from igm_b200.steps._compat import Config
    args = [a for a in sys.argv[1:]]
    ngpu = 0
    if "--gpus" in args:
        k = args.index("--gpus")
        ngpu = int(args[k + 1])
        del args[k:k + 2]
    nstruct = int(args[0]) if args else 1000
    tmp = tempfile.mkdtemp(prefix="igmk_step_")
    dev = torch.device("cuda:0")
    bins = synthetic.genome_bins(200_000)
    chrom_hap, chrom_bead, copy_bead, ci = synthetic.build_index(bins)
    nbead = len(chrom_bead)
    radius = float(synthetic.bead_radius(nbead))
    coords = synthetic.random_walk_coordinates_torch(chrom_bead, copy_bead, nstruct, radius, 20261018, dev).cpu().numpy()
    pop = Population(coords, np.full(nbead, radius, np.float32), chrom_bead, ci, copy_bead)
    t0 = time.perf_counter()
    hss = os.path.join(tmp, "igm-model.hss")
    pop.save_hss(hss)
    pm = synthetic.make_prob_matrix(chrom_hap, seed=20261018)
    hcs = os.path.join(tmp, "in.hcs")
    pm.save_hcs(hcs)
    t_files = time.perf_counter() - t0
    del pop, coords
    out = {"nstruct": nstruct, "nbead": nbead, "write_inputs_s": t_files, "sigmas": [],
           "devices_per_task": ngpu or torch.cuda.device_count()}
    cfg = Config({"parameters": {"workdir": tmp, "tmp_dir": os.path.join(tmp, "tmp")},
                  "optimization": {"structure_output": hss, "iter_corr_knob": 1},
                  "restraints": {"Hi-C": {"input_matrix": hcs, "intra_sigma_list": list(SIGMAS),
                                          "inter_sigma_list": list(SIGMAS), "contact_range": 2.0,
                                          **({"gpu_max_devices": ngpu} if ngpu else {})}},
                  "runtime": {"Hi-C": {}, "opt_iter": 0}})
    for k in range(len(SIGMAS)):
        step = ActivationDistanceStep(cfg)
        ph = {}
        t = time.perf_counter(); step.setup(); ph["setup_s"] = time.perf_counter() - t
        t = time.perf_counter()
        for a in step.argument_list:
            step.task(a, cfg, step.tmp_dir)
        ph["task_s"] = time.perf_counter() - t
        t = time.perf_counter(); step.reduce(); ph["reduce_s"] = time.perf_counter() - t
        with hdf5.open_h5(cfg["runtime"]["Hi-C"]["actdist_file"]) as f:
            ph["records"] = int(len(f["row"]))
        ph["pairs"] = int(step.n_candidate_pairs)
        ph["sigma"] = cfg["runtime"]["Hi-C"]["intra_sigma"]
        import igm_b200.steps  # noqa: F401
        ph["detail"] = {k: round(v, 4) for k, v in sys.modules["igm_b200.steps.ActivationDistanceStep"].LAST_TIMING.items()}
        ph["total_s"] = ph["setup_s"] + ph["task_s"] + ph["reduce_s"]
        ph["pairs_per_s_whole_step"] = ph["pairs"] / ph["total_s"]
        out["sigmas"].append(ph)
        # next sigma (what igm-run does between A/M iterations, bin/igm-run:175-305)
        cfg["runtime"]["Hi-C"].pop("intra_sigma"); cfg["runtime"]["Hi-C"].pop("inter_sigma")
        cfg["runtime"]["opt_iter"] += 1
    out["sweep_pairs"] = int(sum(p["pairs"] for p in out["sigmas"]))
    out["sweep_total_s"] = float(sum(p["total_s"] for p in out["sigmas"]))
    out["sweep_pairs_per_s_whole_step"] = out["sweep_pairs"] / out["sweep_total_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
