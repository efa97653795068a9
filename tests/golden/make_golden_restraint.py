#!/usr/bin/env python
"""Golden vectors for the M-step Hi-C restraint selection, produced by the
reference's OWN intraHiC / interHiC._apply and Particle (igm/restraints/intra_hic.py,
inter_hic.py, igm/model/particle.py) imported through oracle/ref_loader.py.  Build
container only:  python tests/golden/make_golden_restraint.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402
from igm_b200 import synthetic  # noqa: E402


class _Model:
    def __init__(self, particles):
        self.particles = particles
        self.bonds = []

    def addForce(self, f):
        self.bonds.append((int(f.i), int(f.j)))
        return len(self.bonds) - 1


def main():
    ref_loader.install()
    intra = importlib.import_module("igm.restraints.intra_hic").intraHiC
    inter = importlib.import_module("igm.restraints.inter_hic").interHiC
    Particle = importlib.import_module("igm.model.particle").Particle
    out = {}
    pop = synthetic.make_population(2_000_000, 70, seed=8, genome_scale=0.02)
    rng = np.random.default_rng(8)
    nb = pop.nbead
    n_rec = 600
    row = rng.integers(0, nb, n_rec).astype(np.int32)
    col = rng.integers(0, nb, n_rec).astype(np.int32)
    # distances around the actual bead-bead distances so that the test is not trivial,
    # rounded to 4 decimals like the stored actdist column
    s_pick = rng.integers(0, pop.nstruct, n_rec)
    d_true = np.linalg.norm(pop.coordinates[row, s_pick] - pop.coordinates[col, s_pick], axis=1)
    dist = np.array([float("%.4f" % v) for v in d_true * rng.uniform(0.7, 1.3, n_rec)], dtype=np.float32)
    dist[::37] = d_true[::37].astype(np.float32)          # exact ties: norm == dist
    out["coords"], out["radii"], out["chrom"] = pop.coordinates, pop.radii, pop.chrom
    out["row"], out["col"], out["dist"] = row, col, dist
    for kind, cls in (("intra", intra), ("inter", inter)):
        sel = np.zeros((n_rec, pop.nstruct), dtype=bool)
        for s in range(pop.nstruct):
            parts = [Particle(pop.coordinates[b, s], pop.radii[b], Particle.NORMAL) for b in range(nb)]
            r = object.__new__(cls)                 # bypass __init__ (it opens the file with h5py)
            r.contactRange, r.k, r.chrom, r.forceID = 2.0, 1.0, pop.chrom, []
            r.actdist = list(zip(row.tolist(), col.tolist(), dist))
            m = _Model(parts)
            r._apply(m)
            # bonds come out in record order; map back to record indices
            it = iter(m.bonds)
            nxt = next(it, None)
            for k in range(n_rec):
                if nxt is not None and nxt == (int(row[k]), int(col[k])) and _hit(parts, row[k], col[k], dist[k], pop.chrom, kind):
                    sel[k, s] = True
                    nxt = next(it, None)
            assert nxt is None
        out["sel_" + kind] = np.packbits(sel, axis=1)
        print(kind, int(sel.sum()), "of", sel.size)
    np.savez_compressed(os.path.join(HERE, "restraint_small.npz"), **out)


def _hit(parts, i, j, d, chrom, kind):
    same = chrom[i] == chrom[j]
    return ((parts[i] - parts[j]) <= d) and (same if kind == "intra" else not same)


if __name__ == "__main__":
    main()
