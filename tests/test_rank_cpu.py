"""Rank-matching oracle (FISH / polymer assignment, row f4) pinned to the reference's own
get_rad_dists / get_pair_dists / get_min_max_and_idx / get_polymer_dists
(golden vectors from tests/golden/make_golden_rank.py)."""
import os

import numpy as np
import pytest

from oracle import rank_oracle as ro

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "rank_small.npz")
POPS = os.path.join(os.path.dirname(__file__), "golden", "damid_small.npz")


def load_pop(name):
    g = np.load(POPS)
    ptr, beads = g[name + "_copy_ptr"], g[name + "_copy_beads"]
    ci = [[int(b) for b in beads[ptr[i]:ptr[i + 1]]] for i in range(len(ptr) - 1)]
    return g[name + "_coords"], g[name + "_radii"], g[name + "_chrom"], ptr, beads, ci


@pytest.mark.parametrize("name", ["n37", "n100", "n257"])
def test_oracle_matches_reference_golden(name):
    g = np.load(GOLDEN)
    crd, _, _, _, _, ci = load_pop(name)
    for k, i in enumerate(g[name + "_probes"]):
        mind, maxd, imin, imax = ro.min_max_and_idx(ro.radial_values(crd, ci[i]))
        assert np.array_equal(mind.astype(np.float32), g[name + "_rad_min"][k])
        assert np.array_equal(maxd.astype(np.float32), g[name + "_rad_max"][k])
        assert np.array_equal(imin, g[name + "_rad_imin"][k]) and np.array_equal(imax, g[name + "_rad_imax"][k])
    for k, (i, j) in enumerate(g[name + "_pairs"]):
        mind, _, imin, _ = ro.min_max_and_idx(ro.pair_values(crd, ci[i], ci[j]))
        assert np.array_equal(mind.astype(np.float32), g[name + "_pair_min"][k])
        assert np.array_equal(imin, g[name + "_pair_imin"][k])
    for i in range(crd.shape[0] - 1):
        d, idx = ro.polymer_dists(crd, i)
        assert d.dtype == np.float32 and np.array_equal(d, g[name + "_poly_dist"][i])
        assert np.array_equal(idx, g[name + "_poly_idx"][i])


def test_norm_model_equals_numpy():
    rng = np.random.default_rng(5)
    d = (rng.standard_normal((50000, 3)) * 3000).astype(np.float32)
    assert np.array_equal(ro.norm32(d), np.linalg.norm(d, axis=1))


def test_ties_are_ranked_by_structure_index():
    v = np.array([3.0, 1.0, 3.0, 0.5, 1.0, 3.0])
    assert ro.stable_rank(v).tolist() == [3, 1, 4, 0, 2, 5]
    r = np.argsort(np.argsort(v))               # NumPy's default order inside tie groups may differ
    assert ro.same_up_to_ties(v, r, ro.stable_rank(v))
    assert not ro.same_up_to_ties(v, np.array([0, 1, 2, 3, 4, 5]), ro.stable_rank(v))


def test_fish_and_polymer_step_host_logic(tmp_path):
    """setup()/skip() of the two drop-ins without a GPU: batch layout (FishAssignmentStep.py:129-155,
    PolymerAssignmentStep.py:70-82), names, directories."""
    from igm_b200 import hdf5, synthetic
    from igm_b200.steps import FishAssignmentStep, PolymerAssignmentStep
    from igm_b200.steps._compat import Config
    pop = synthetic.make_population(2_000_000, 8, seed=3, genome_scale=0.02)
    hss = str(tmp_path / "pop.hss")
    pop.save_hss(hss)
    fish_in = str(tmp_path / "fish.h5")
    hdf5.write_h5(fish_in, {"pairs": np.array([[0, 1], [2, 3], [4, 5]], np.int32),
                            "probes": np.array([7, 8, 9, 10, 11], np.int32)})
    cfg = Config({"parameters": {"workdir": str(tmp_path), "tmp_dir": str(tmp_path / "tmp")},
                  "optimization": {"structure_output": hss},
                  "restraints": {"FISH": {"input_fish": fish_in, "tol_list": [100.0], "batch_size": 2},
                                 "polymer": {"polymer_file": "x", "assignment_file": "polymer_assignment.h5"}},
                  "runtime": {"FISH": {}, "polymer": {}}})
    f = FishAssignmentStep(cfg)
    assert f.name() == "FishAssignmentStep (tol=100.00, iter=N/A)" and cfg["runtime"]["FISH"]["tol_list"] == []
    f.setup()
    assert [(b[0], b[1], np.asarray(b[2]).tolist()) for b in f.argument_list] == [
        (0, "pair", [[0, 1], [2, 3]]), (1, "pair", [[4, 5]]),
        (2, "probe", [7, 8]), (3, "probe", [9, 10]), (4, "probe", [11])]
    f.skip()
    assert cfg["runtime"]["FISH"]["fish_assignment_file"] == str(tmp_path / "tmp" / "fish_actdist" / "fish_assignment.h5")
    p = PolymerAssignmentStep(cfg)
    assert p.name() == "PolymerAssignmentStep (iter=N/A)"
    p.setup()
    assert len(p.argument_list) == 1 and list(p.argument_list[0][1]) == list(range(pop.nbead - 1))
    p.skip()
    assert cfg["runtime"]["polymer"]["assignment_file"] == str(tmp_path / "tmp" / "poly_actdist" / "polymer_assignment.h5")
