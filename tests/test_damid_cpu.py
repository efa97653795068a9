"""DamID oracle pinned to the reference's own get_damid_actdist_I (golden vectors
from tests/golden/make_golden_damid.py); host logic of the Step drop-in."""
import hashlib
import os

import numpy as np
import pytest

from oracle import damid_oracle as do

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "damid_small.npz")


def load_case(g, name):
    ptr, beads = g[name + "_copy_ptr"], g[name + "_copy_beads"]
    ci = {i: [int(b) for b in beads[ptr[i]:ptr[i + 1]]] for i in range(len(ptr) - 1)}
    return (g[name + "_coords"], g[name + "_radii"], g[name + "_chrom"], ptr, beads, ci,
            g[name + "_params"], float(g[name + "_nucleus_radius"]))


@pytest.mark.parametrize("name", ["n37", "n100", "n257"])
@pytest.mark.parametrize("it_corr", [0, 1])
def test_oracle_matches_reference_golden(name, it_corr):
    g = np.load(GOLDEN)
    coords, radii, chrom, ptr, beads, ci, params, rad = load_case(g, name)
    recs, dets = do.run_loci(params[:, 0], params[:, 1], params[:, 2], coords, radii, ci, it_corr, 0.05, rad)
    key = "%s_it%d" % (name, it_corr)
    assert np.array_equal(np.array([r[0] for r in recs], np.int32), g[key + "_loc"])
    ad = np.array([float(r[1]) for r in recs])
    p = np.array([float(r[2]) for r in recs])
    assert np.array_equal(ad.view(np.uint64), g[key + "_ad"].view(np.uint64))
    assert np.array_equal(p.view(np.uint64), g[key + "_p"].view(np.uint64))
    assert hashlib.sha256(do.task_text(recs).encode()).hexdigest() == str(g[key + "_sha"])
    # cases the golden set must exercise: p <= 0 (distance 2), plast >= 1, haploid loci
    assert any(d["o"] < 0 for d in dets) and any(d["ncopies"] == 1 for d in dets)


def test_text_roundtrip5():
    v = np.array([0.000005, 0.000015, 1.234565, 2.0, 0.999995, 1e-9])
    assert np.array_equal(do.text_roundtrip5(v), np.array([float("%.5f" % x) for x in v], np.float32))


def test_expand_damid_records_order():
    """Record expansion (one record per copy, locus order) without a GPU."""
    from igm_b200.engine import ActdistEngine
    from igm_b200._lib import PAIR_RESULT_DTYPE
    ptr = np.array([0, 2, 3, 5])
    beads = np.array([0, 3, 1, 2, 4])
    res = np.zeros(2, PAIR_RESULT_DTYPE)
    res["nrec"] = [2, 1]
    res["dist"] = [0.5, 2.0]
    res["prob"] = [0.25, 0.0]
    loc, dist, prob = ActdistEngine.expand_damid_records(None, np.array([2, 1]), res, ptr, beads)
    assert loc.tolist() == [2, 4, 1] and dist.tolist() == [0.5, 0.5, 2.0] and prob.tolist() == [0.25, 0.25, 0.0]


def test_nucl_damid_step_host_logic(tmp_path):
    """Host side of the nuclear-body twin (NuclDamidActivationDistanceStep.py:120-361): config keys,
    name string, sigma pop, batch files, and the setup/skip default-directory mismatch (:172 vs :356)."""
    from igm_b200.steps import NuclDamidActivationDistanceStep
    from igm_b200.steps._compat import Config
    prof = str(tmp_path / "p.txt")
    np.savetxt(prof, np.array([0.5, 0.05, 0.31, 0.9, 0.2], np.float32))
    cfg = Config({"parameters": {"workdir": str(tmp_path), "tmp_dir": str(tmp_path / "tmp")},
                  "optimization": {"structure_output": "unused.hss", "iter_corr_knob": 0},
                  "restraints": {"nuclDamID": {"input_profile": prof, "sigma_list": [0.3, 0.1], "batch_size": 2}},
                  "runtime": {"nuclDamID": {}, "opt_iter": 4}})
    step = NuclDamidActivationDistanceStep(cfg)
    assert step.name() == "NuclDamidActivationDistanceStep (sigma=30.00%, iter=4)"
    assert cfg["runtime"]["nuclDamID"]["sigma"] == 0.3 and cfg["runtime"]["nuclDamID"]["sigma_list"] == [0.1]
    assert "DamID" not in cfg["runtime"]
    step.setup()
    assert step.tmp_dir == str(tmp_path / "tmp" / "nucldamid_actdist")
    assert list(step.argument_list) == [0, 1]
    b0 = np.load(os.path.join(step.tmp_dir, "0.damid.in.npy"))
    b1 = np.load(os.path.join(step.tmp_dir, "1.damid.in.npy"))
    assert b0.dtype == np.float32 and b0[:, 0].tolist() == [0.0, 2.0] and b1[:, 0].tolist() == [3.0]
    step.skip()
    assert cfg["runtime"]["nuclDamID"]["damid_actdist_file"] == str(tmp_path / "tmp" / "damid_actdist" / "damid_actdist.hdf5")
