#!/usr/bin/env python
"""Golden vectors for the DamID activation distance, produced by the reference's
OWN ``get_damid_actdist_I`` (igm/steps/DamidActivationDistanceStep.py:375-470)
imported through oracle/ref_loader.py.  Runs only in the build container
(/root/reference present):  python tests/golden/make_golden_damid.py
"""
import contextlib
import hashlib
import importlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402
from igm_b200 import synthetic  # noqa: E402


def main():
    ref_loader.install()
    mod = importlib.import_module("igm.steps.DamidActivationDistanceStep")
    fn = mod.get_damid_actdist_I
    fmt = mod.damid_actdist_fmt_str
    out = {}
    for name, nstruct, seed in (("n37", 37, 1), ("n100", 100, 2), ("n257", 257, 3)):
        pop = synthetic.make_population(2_000_000, nstruct, seed=seed, genome_scale=0.02)
        hss = ref_loader.FakeHss(pop.coordinates, pop.radii, pop.chrom, pop.copy_index.to_dict())
        rng = np.random.default_rng(seed)
        nh = pop.n_hap
        loci = np.arange(nh)
        p_exp = rng.uniform(0.0, 1.0, nh)
        p_exp[::7] = 1.0
        p_exp[3::11] = 0.0
        plast = np.where(rng.random(nh) < 0.5, 0.0, rng.uniform(0, 0.9, nh))
        plast[5::13] = 1.0
        params = np.array(list(zip(loci, p_exp, plast)), dtype=np.float32)
        # nucleus radius chosen so that a sizeable fraction of beads is "in contact"
        rad = float(np.sqrt(np.quantile(np.sum(np.square(pop.coordinates), axis=2), 0.7))) / 0.95 + float(pop.radii[0])
        out[name + "_coords"] = pop.coordinates
        out[name + "_radii"] = pop.radii
        out[name + "_chrom"] = pop.chrom
        out[name + "_copy_ptr"] = pop.copy_index.ptr
        out[name + "_copy_beads"] = pop.copy_index.beads
        out[name + "_params"] = params
        out[name + "_nucleus_radius"] = np.float64(rad)
        for it_corr in (0, 1):
            recs = []
            with contextlib.redirect_stdout(io.StringIO()):      # the reference prints per locus
                for I, pe, pl in params:
                    recs += fn(int(I), pe, pl, hss, it_corr, contact_range=0.05, shape="sphere",
                               nucleus_param=rad)
            key = "%s_it%d" % (name, it_corr)
            out[key + "_loc"] = np.array([r[0] for r in recs], np.int32)
            out[key + "_ad"] = np.array([float(r[1]) for r in recs], np.float64)
            out[key + "_p"] = np.array([float(r[2]) for r in recs], np.float64)
            text = "\n".join([fmt % x for x in recs])
            out[key + "_sha"] = np.array(hashlib.sha256(text.encode()).hexdigest())
            print(key, len(recs), out[key + "_sha"])
    np.savez_compressed(os.path.join(HERE, "damid_small.npz"), **out)


if __name__ == "__main__":
    main()
