#!/usr/bin/env python
"""Golden vectors for the rank matching of the FISH and polymer assignment steps, produced
by the reference's OWN functions (igm/steps/FishAssignmentStep.py:23-79 get_pair_dists,
get_rad_dists, get_min_max_and_idx; igm/steps/PolymerAssignmentStep.py:24-33
get_polymer_dists) imported through oracle/ref_loader.py.  Populations are those of
tests/golden/damid_small.npz.  Runs only in the build container (/root/reference present):
    python tests/golden/make_golden_rank.py
Pairs are restricted to single-copy loci: with more combinations the reference's
get_pair_dists returns uninitialised rows.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402


def main():
    ref_loader.install()
    fish = importlib.import_module("igm.steps.FishAssignmentStep")
    poly = importlib.import_module("igm.steps.PolymerAssignmentStep")
    g = np.load(os.path.join(HERE, "damid_small.npz"))
    out = {}
    for name in ("n37", "n100", "n257"):
        crd = g[name + "_coords"]
        ptr, beads = g[name + "_copy_ptr"], g[name + "_copy_beads"]
        nbead, nconf = crd.shape[0], crd.shape[1]
        ci = [[int(b) for b in beads[ptr[i]:ptr[i + 1]]] for i in range(len(ptr) - 1)]
        dip = [i for i, c in enumerate(ci) if len(c) == 2]
        hap = [i for i, c in enumerate(ci) if len(c) == 1]
        # probes: every diploid locus (get_rad_dists needs two copies)
        rmin, rmax, imin, imax = [], [], [], []
        for i in dip:
            d = fish.get_rad_dists(ci[i], nbead, nconf, crd)
            a, b, c, e = fish.get_min_max_and_idx(d)
            rmin.append(a); rmax.append(b); imin.append(c); imax.append(e)
        out[name + "_probes"] = np.array(dip, np.int32)
        out[name + "_rad_min"] = np.array(rmin); out[name + "_rad_max"] = np.array(rmax)
        out[name + "_rad_imin"] = np.array(imin, np.int32); out[name + "_rad_imax"] = np.array(imax, np.int32)
        # pairs of single-copy loci
        rng = np.random.default_rng(len(hap))
        pairs = [(hap[a], hap[b]) for a, b in rng.integers(0, len(hap), (40, 2)) if a != b]
        pmin, pimin = [], []
        for i, j in pairs:
            d = fish.get_pair_dists(ci[i], ci[j], nbead, nconf, crd)
            assert d.shape == (1, nconf)
            a, b, c, e = fish.get_min_max_and_idx(d)
            assert np.array_equal(a, b) and np.array_equal(c, e)
            pmin.append(a); pimin.append(c)
        out[name + "_pairs"] = np.array(pairs, np.int32)
        out[name + "_pair_min"] = np.array(pmin); out[name + "_pair_imin"] = np.array(pimin, np.int32)
        # polymer bonds (i, i+1) over all beads
        pd, pi = [], []
        for i in range(nbead - 1):
            d, idx = poly.get_polymer_dists(i, crd)
            pd.append(d); pi.append(idx)
        out[name + "_poly_dist"] = np.array(pd); out[name + "_poly_idx"] = np.array(pi, np.int32)
        nties = sum(len(np.unique(v)) != len(v) for v in rmin + rmax + pmin + pd)
        print(name, len(dip), "probes", len(pairs), "pairs", nbead - 1, "bonds; rows with ties:", nties,
              out[name + "_poly_dist"].dtype, out[name + "_rad_min"].dtype)
        assert nties == 0, "regenerate with another population: ties make the reference's ranks arbitrary"
    # the float64 arrays of the FISH helpers only ever hold float32 values: store them as such
    for k in list(out):
        if out[k].dtype == np.float64:
            assert np.array_equal(out[k].astype(np.float32).astype(np.float64), out[k])
            out[k] = out[k].astype(np.float32)
        elif k.endswith(("_imin", "_imax", "_idx")):
            out[k] = out[k].astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "rank_small.npz"), **out)


if __name__ == "__main__":
    main()
