#!/usr/bin/env python
"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports the reference's own ``get_actdist`` (LB: igm/steps/
ActivationDistanceStep.py:336, GP: igm/steps/GP_activation.py:317,
igm/utils/actdist.py:15) through ``oracle/ref_loader.py`` and writes

* ``demo_subset.npz``     - a 96-bin slice of the shipped demo population
  (demo/demo_sample_outputs/igm-model.hss.T) with reference outputs for a set
  of pairs in every mode (LB/GP x it_corr 0/1), incl. X/Y haploid bins.
* ``synth_small.npz``     - seeded synthetic populations with ragged sizes
  (N = 1, 3, 37, 100, 257) and their reference outputs.
* ``demo_full.json``      - sha256 of the reference's '%d.out.tmp' text for the
  whole demo population at every sigma of demo/config_file.json, plus pair /
  record counts (checked against SURVEY.md 8c).
* ``demo_sigma005.npz``   - per-pair reference details (d2 bits, count, o, p)
  for all 14 994 candidates at sigma = 0.05, LB and GP.

It also copies the two demo HDF5 files into git-ignored ``oracle/_ref/demo/``
so that the full-demo parity tests can run on the GPU box, where
/root/reference does not exist.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import actdist_oracle as orc  # noqa: E402
from igm_b200.population import Population, ProbMatrix, CopyIndex  # noqa: E402
from igm_b200 import synthetic  # noqa: E402

REF = ref_loader.REFERENCE_ROOT
DEMO_HSS = os.path.join(REF, "demo/demo_sample_outputs/igm-model.hss.T")
DEMO_HCS = os.path.join(REF, "demo/WTC11_HiC_2Mb.hcs")


def run_reference(fn, pop, ii, jj, pw, pl, it_corr, with_it_corr_arg=True, contact_range=2.0):
    """Loop of task() (ActivationDistanceStep.py:215-222) over the reference fn."""
    hss = ref_loader.FakeHss(pop.coordinates, pop.radii, pop.chrom_hap(), pop.copy_index)
    recs = []
    nrec = np.zeros(len(ii), np.int32)
    for k, (a, b, w, l) in enumerate(zip(ii, jj, pw, pl)):
        if with_it_corr_arg:
            r = fn(int(a), int(b), np.float64(w), np.float64(l), hss, it_corr,
                   contactRange=contact_range)
        else:
            r = fn(int(a), int(b), np.float64(w), np.float64(l), hss, contactRange=contact_range)
        nrec[k] = len(r)
        recs.extend(r)
    return recs, nrec


def rec_arrays(recs):
    if not recs:
        return (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), np.zeros(0))
    return (np.array([r[0] for r in recs], np.int32), np.array([r[1] for r in recs], np.int32),
            np.array([float(r[2]) for r in recs], np.float64),
            np.array([float(r[3]) for r in recs], np.float64))


def pack_case(out, prefix, ref, pop, ii, jj, pw, pl, contact_range=2.0):
    fmt = ref["actdist_fmt_str"]
    for mode, fn in (("lb", ref["lb_get_actdist"]), ("gp", ref["gp_get_actdist"])):
        for it_corr in (0, 1):
            recs, nrec = run_reference(fn, pop, ii, jj, pw, pl, it_corr, contact_range=contact_range)
            row, col, ad, p = rec_arrays(recs)
            key = "%s_%s_it%d" % (prefix, mode, it_corr)
            out[key + "_row"] = row
            out[key + "_col"] = col
            out[key + "_ad"] = ad
            out[key + "_p"] = p
            out[key + "_nrec"] = nrec
            text = "\n".join([fmt % x for x in recs])
            out[key + "_sha"] = np.array(hashlib.sha256(text.encode()).hexdigest())


def save_pop(out, prefix, pop, ii, jj, pw, pl):
    out[prefix + "_coords"] = pop.coordinates
    out[prefix + "_radii"] = pop.radii
    out[prefix + "_chrom"] = pop.chrom
    out[prefix + "_copy_ptr"] = pop.copy_index.ptr
    out[prefix + "_copy_beads"] = pop.copy_index.beads
    out[prefix + "_i"] = np.asarray(ii, np.int32)
    out[prefix + "_j"] = np.asarray(jj, np.int32)
    out[prefix + "_pwish"] = np.asarray(pw, np.float64)
    out[prefix + "_plast"] = np.asarray(pl, np.float64)


def demo_subset(ref, pop, pm):
    """96 haploid bins: 40 from chr1, 30 from chr7, 16 from X, 10 from Y."""
    ch = pop.chrom_hap()
    sel = np.concatenate([np.nonzero(ch == 0)[0][:40], np.nonzero(ch == 6)[0][:30],
                          np.nonzero(ch == 22)[0][:16], np.nonzero(ch == 23)[0][:10]])
    n_sub = len(sel)
    nc = pop.copy_index.ncopies()[sel]
    dip = sel[nc == 2]
    # new bead order: copy-0 of every selected bin, then copy-1 of diploid ones
    beads0 = np.array([pop.copy_index[s][0] for s in sel])
    beads1 = np.array([pop.copy_index[s][1] for s in dip])
    beads = np.concatenate([beads0, beads1])
    assert np.all(nc[:len(dip)] == 2) and np.all(nc[len(dip):] == 1)
    sub = Population(pop.coordinates[beads], pop.radii[beads], pop.chrom[beads],
                     CopyIndex.diploid(n_sub, len(dip)))
    rng = np.random.default_rng(7)
    # all pairs among the first 12 bins + random pairs over everything
    ii, jj = np.triu_indices(12, 1)
    ri = rng.integers(0, n_sub, 400)
    rj = rng.integers(0, n_sub, 400)
    k = ri < rj
    ii = np.concatenate([ii, ri[k]])
    jj = np.concatenate([jj, rj[k]])
    # pwish: real demo values where stored, else random; include 1.0 and tiny values
    pw = np.empty(len(ii), np.float32)
    for t, (a, b) in enumerate(zip(sel[ii], sel[jj])):
        lo, hi = pm.indptr[a], pm.indptr[a + 1]
        pos = np.searchsorted(pm.indices[lo:hi], b)
        if pos < hi - lo and pm.indices[lo + pos] == b:
            pw[t] = pm.data[lo + pos]
        else:
            pw[t] = np.float32(rng.uniform(0.001, 0.3))
    pw[::17] = np.float32(1.0)
    pw[5::23] = np.float32(0.00004)
    pl = np.where(rng.random(len(ii)) < 0.5, 0.0,
                  orc.text_roundtrip(rng.uniform(0, 0.4, len(ii))).astype(np.float64))
    pl[3::29] = 1.0
    return sub, ii.astype(np.int32), jj.astype(np.int32), pw.astype(np.float64), pl


def synth_cases():
    cases = []
    for n, scale, seed in ((1, 0.012, 11), (3, 0.012, 12), (37, 0.015, 13),
                           (100, 0.02, 14), (257, 0.012, 15)):
        pop = synthetic.make_population(2_000_000, n, seed=seed, genome_scale=scale)
        rng = np.random.default_rng(seed + 100)
        nh = pop.n_hap
        m = 160
        ri = rng.integers(0, nh, m)
        rj = rng.integers(0, nh, m)
        k = ri < rj
        ii, jj = ri[k], rj[k]
        # adjacent bins (p = 1 neighbours) as well
        adj = np.arange(0, nh - 1, 3)
        ii = np.concatenate([ii, adj])
        jj = np.concatenate([jj, adj + 1])
        pw = rng.uniform(0.0005, 1.0, len(ii)).astype(np.float32)
        pw[::5] = np.float32(1.0)
        pw[1::7] = np.float32(0.01)
        pl = np.where(rng.random(len(ii)) < 0.4, 0.0,
                      orc.text_roundtrip(rng.uniform(0, 0.9, len(ii))).astype(np.float64))
        cases.append(("n%d" % n, pop, ii.astype(np.int32), jj.astype(np.int32),
                      pw.astype(np.float64), pl))
    return cases


def main():
    ref = ref_loader.load_reference()
    pop = Population.from_hss(DEMO_HSS)
    pm = ProbMatrix.from_hcs(DEMO_HCS)
    os.makedirs(os.path.join(ROOT, "oracle/_ref/demo"), exist_ok=True)
    for src in (DEMO_HSS, DEMO_HCS):
        dst = os.path.join(ROOT, "oracle/_ref/demo", os.path.basename(src))
        if not os.path.exists(dst):
            shutil.copyfile(src, dst)

    # 1. demo subset
    out = {}
    sub, ii, jj, pw, pl = demo_subset(ref, pop, pm)
    save_pop(out, "demo", sub, ii, jj, pw, pl)
    pack_case(out, "demo", ref, sub, ii, jj, pw, pl)
    # utils/actdist.get_actdist (no it_corr argument; always corrects)
    recs, nrec = run_reference(ref["utils_get_actdist"], sub, ii, jj, pw, pl, 1, with_it_corr_arg=False)
    row, col, ad, p = rec_arrays(recs)
    out["demo_utils_ad"] = ad
    out["demo_utils_p"] = p
    np.savez_compressed(os.path.join(HERE, "demo_subset.npz"), **out)
    print("demo_subset: %d pairs, %d beads" % (len(ii), sub.nbead))

    # 2. synthetic ragged cases (+ a contact_range != 2 variant)
    out = {}
    names = []
    for name, spop, ii, jj, pw, pl in synth_cases():
        names.append(name)
        save_pop(out, name, spop, ii, jj, pw, pl)
        pack_case(out, name, ref, spop, ii, jj, pw, pl)
        print("synth %s: %d pairs, %d beads" % (name, len(ii), spop.nbead))
    name, spop, ii, jj, pw, pl = synth_cases()[3]
    pack_case(out, "n100cr35", ref, spop, ii, jj, pw, pl, contact_range=3.5)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "synth_small.npz"), **out)

    # 3. full demo population, sigma sweep (demo/config_file.json:35-50)
    summary = {"source": "reference get_actdist run by tests/golden/make_golden.py",
               "numpy": np.__version__, "sigmas": {}}
    fmt = ref["actdist_fmt_str"]
    d005 = {}
    for sigma in (1.0, 0.2, 0.1, 0.05, 0.02, 0.01):
        # float64 comparison, as in the SURVEY probe
        ci, cj, cp = orc.select_candidates(pm.indptr, pm.indices, pm.data, pm.chrom,
                                           sigma, sigma, compare_dtype="float64")
        n32 = len(orc.select_candidates(pm.indptr, pm.indices, pm.data, pm.chrom,
                                        sigma, sigma, compare_dtype="float32")[0])
        entry = {"pairs_f64_compare": int(len(ci)), "pairs_f32_compare": int(n32)}
        modes = [("lb", ref["lb_get_actdist"])]
        if sigma >= 0.02:
            modes.append(("gp", ref["gp_get_actdist"]))
        for mode, fn in modes:
            recs, nrec = run_reference(fn, pop, ci, cj, cp, np.zeros(len(ci)), 0)
            text = "\n".join([fmt % x for x in recs])
            entry[mode] = {"records": len(recs),
                           "sha256": hashlib.sha256(text.encode()).hexdigest()}
            print("sigma %g %s: %d pairs %d records %s" % (
                sigma, mode, len(ci), len(recs), entry[mode]["sha256"][:16]))
            if sigma == 0.05:
                row, col, ad, p = rec_arrays(recs)
                first = np.concatenate([[0], np.cumsum(nrec)[:-1]])
                d005[mode + "_ad"] = ad[first]
                d005[mode + "_p"] = p[first]
                d005[mode + "_nrec"] = nrec
                d005["i"], d005["j"], d005["pwish"] = ci, cj, cp
        summary["sigmas"]["%g" % sigma] = entry
    np.savez_compressed(os.path.join(HERE, "demo_sigma005.npz"), **d005)
    with open(os.path.join(HERE, "demo_full.json"), "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
