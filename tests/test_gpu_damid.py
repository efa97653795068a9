"""GPU parity of the DamID activation distance (next row f1) through the C ABI:
against the golden vectors of the reference's own function and the oracle."""
import hashlib
import os

import numpy as np
import pytest

from oracle import damid_oracle as do
from tests.test_damid_cpu import GOLDEN, load_case

pytestmark = pytest.mark.gpu


def _pop(coords, radii, chrom, ptr, beads):
    from igm_b200.population import CopyIndex, Population
    return Population(coords, radii, chrom, CopyIndex(ptr, beads), None)


@pytest.mark.parametrize("name", ["n37", "n100", "n257"])
@pytest.mark.parametrize("it_corr", [0, 1])
def test_damid_golden(name, it_corr):
    from igm_b200.engine import ActdistEngine
    g = np.load(GOLDEN)
    coords, radii, chrom, ptr, beads, ci, params, rad = load_case(g, name)
    key = "%s_it%d" % (name, it_corr)
    with ActdistEngine(_pop(coords, radii, chrom, ptr, beads), 0) as eng:
        loci = params[:, 0].astype(np.int32)
        res = eng.damid_actdist(loci, params[:, 1], params[:, 2], rad, 0.05, it_corr)
        loc, dist, prob = eng.expand_damid_records(loci, res, ptr, beads)
    assert np.array_equal(loc, g[key + "_loc"])
    assert np.array_equal(dist.view(np.uint32), do.text_roundtrip5(g[key + "_ad"]).view(np.uint32))
    assert np.array_equal(prob.view(np.uint32), do.text_roundtrip5(g[key + "_p"]).view(np.uint32))
    # full-precision fields and the text of '%d.out.tmp'
    _, dets = do.run_loci(params[:, 0], params[:, 1], params[:, 2], coords, radii, ci, it_corr, 0.05, rad)
    assert np.array_equal(res["o"], np.array([d["o"] for d in dets]))
    assert np.array_equal(res["p"], np.array([d["p"] for d in dets]))
    assert np.array_equal(res["contact_count"], np.array([d["contact_count"] for d in dets]))
    sel = res["o"] >= 0
    assert np.array_equal(res["d2_sel_bits"][sel], np.array([d["s_bits"] for d in dets], np.uint32)[sel])
    rep = np.repeat(np.arange(len(loci)), res["nrec"])
    den = np.array([d["denom"] for d in dets])
    ad = np.where(sel, np.sqrt(res["d2_sel_bits"].view(np.float32).astype(np.float64) / den), 2.0)
    text = do.task_text(list(zip(loc.tolist(), ad[rep].tolist(), res["p"][rep].tolist())))
    # '%.5f' % 2 (an int in the reference) prints like 2.0
    assert hashlib.sha256(text.encode()).hexdigest() == str(g[key + "_sha"])


@pytest.mark.parametrize("nstruct", [5, 1000, 2600])
def test_damid_oracle_sizes(nstruct):
    """Warp and CTA groups, ragged populations, every locus."""
    from igm_b200 import synthetic
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, nstruct, seed=nstruct, genome_scale=0.01)
    rng = np.random.default_rng(nstruct)
    nh = pop.n_hap
    loci = np.arange(nh)
    pe = rng.uniform(0, 1, nh).astype(np.float32)
    pl = np.where(rng.random(nh) < 0.5, 0, rng.uniform(0, 1, nh)).astype(np.float32)
    rad = float(np.sqrt(np.quantile(np.sum(np.square(pop.coordinates), axis=2), 0.6))) / 0.95 + float(pop.radii[0])
    with ActdistEngine(pop, 0) as eng:
        for it_corr in (0, 1):
            res = eng.damid_actdist(loci, pe, pl, rad, 0.05, it_corr)
            _, dets = do.run_loci(loci, pe, pl, pop.coordinates, pop.radii, pop.copy_index, it_corr, 0.05, rad)
            assert np.array_equal(res["o"], np.array([d["o"] for d in dets]))
            assert np.array_equal(res["p"], np.array([d["p"] for d in dets]))
            assert np.array_equal(res["contact_count"], np.array([d["contact_count"] for d in dets]))
            sel = res["o"] >= 0
            assert np.array_equal(res["d2_sel_bits"][sel], np.array([d["s_bits"] for d in dets], np.uint32)[sel])


@pytest.mark.parametrize("variant", ["DamID", "nuclDamID"])
def test_damid_step_end_to_end(tmp_path, variant):
    """Lamina step and its nuclear-body twin (NuclDamidActivationDistanceStep.py) through
    Step.run on files; both must give the oracle's records."""
    from igm_b200 import hdf5, synthetic
    from igm_b200 import steps
    from igm_b200.steps._compat import Config
    body = {"DamID": "envelope", "nuclDamID": "nucleolus"}[variant]
    cls = {"DamID": steps.DamidActivationDistanceStep, "nuclDamID": steps.NuclDamidActivationDistanceStep}[variant]
    pop = synthetic.make_population(2_000_000, 80, seed=12, genome_scale=0.02)
    hss = str(tmp_path / "pop.hss")
    pop.save_hss(hss)
    rng = np.random.default_rng(3)
    profile = rng.uniform(0, 1, pop.n_hap).astype(np.float32)
    prof = str(tmp_path / "damid.txt")
    np.savetxt(prof, profile)
    rad = float(np.sqrt(np.quantile(np.sum(np.square(pop.coordinates), axis=2), 0.6))) / 0.95 + float(pop.radii[0])
    cfg = Config({"parameters": {"workdir": str(tmp_path), "tmp_dir": str(tmp_path / "tmp")},
                  "optimization": {"structure_output": hss, "iter_corr_knob": 1},
                  "model": {"restraints": {body: {"nucleus_shape": "sphere", "nucleus_radius": rad}}},
                  "restraints": {variant: {"input_profile": prof, "sigma_list": [0.3, 0.1], "contact_range": 0.05,
                                           "batch_size": 17, "keep_temporary_files": True}},
                  "runtime": {variant: {}, "opt_iter": 1}})
    step = cls(cfg)
    assert step.name() == cls.__name__ + " (sigma=30.00%, iter=1)"
    step.run()
    out = cfg["runtime"][variant]["damid_actdist_file"]
    assert os.path.basename(os.path.dirname(out)) == {"DamID": "damid_actdist", "nuclDamID": "nucldamid_actdist"}[variant]
    with hdf5.open_h5(out) as f:
        loc, dist, prob = np.asarray(f["loc"][()]), np.asarray(f["dist"][()]), np.asarray(f["prob"][()])
    prof32 = np.loadtxt(prof, dtype="float32")
    ii = np.where(prof32 >= 0.3)[0]
    recs, _ = do.run_loci(ii, prof32[ii], np.zeros(len(ii)), pop.coordinates, pop.radii, pop.copy_index, 1, 0.05, rad)
    assert np.array_equal(loc, np.array([r[0] for r in recs], np.int32))
    assert np.array_equal(dist, do.text_roundtrip5([r[1] for r in recs]))
    assert np.array_equal(prob, do.text_roundtrip5([r[2] for r in recs]))
    assert loc.dtype == np.int32 and dist.dtype == np.float32 and prob.dtype == np.float32
