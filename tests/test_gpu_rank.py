"""GPU parity of the rank matching (K5; FISH and polymer assignment steps, next row f4)
through the C ABI: against the golden vectors of the reference's own helper functions,
against the oracle at other sizes and with ties, and through the Step drop-ins on files."""
import os

import numpy as np
import pytest

from oracle import rank_oracle as ro
from tests.test_rank_cpu import GOLDEN, load_pop

pytestmark = pytest.mark.gpu


def _pop(coords, radii, chrom, ptr, beads):
    from igm_b200.population import CopyIndex, Population
    return Population(coords, radii, chrom, CopyIndex(ptr, beads), None)


def _copies2(ci, loci):
    return np.array([[ci[i][0], ci[i][1] if len(ci[i]) > 1 else -1] for i in loci], np.int32)


@pytest.mark.parametrize("name", ["n37", "n100", "n257"])
def test_rank_golden(name):
    from igm_b200.engine import ActdistEngine
    g = np.load(GOLDEN)
    crd, radii, chrom, ptr, beads, ci = load_pop(name)
    n = crd.shape[1]
    tgt = np.sort(np.random.default_rng(1).uniform(0, 4000, n)).astype(np.float32)
    with ActdistEngine(_pop(crd, radii, chrom, ptr, beads), 0) as eng:
        a = _copies2(ci, g[name + "_probes"])
        lo = eng.rank_match(a, None, "min", tgt)
        hi = eng.rank_match(a, None, "max", tgt)
        assert np.array_equal(lo["value"].view(np.uint32), g[name + "_rad_min"].view(np.uint32))
        assert np.array_equal(hi["value"].view(np.uint32), g[name + "_rad_max"].view(np.uint32))
        assert np.array_equal(lo["rank"], g[name + "_rad_imin"]) and np.array_equal(hi["rank"], g[name + "_rad_imax"])
        assert np.array_equal(lo["matched"], tgt[g[name + "_rad_imin"].astype(np.int64)])
        pairs = g[name + "_pairs"]
        pr = eng.rank_match(_copies2(ci, pairs[:, 0]), _copies2(ci, pairs[:, 1]), "min")
        assert np.array_equal(pr["value"].view(np.uint32), g[name + "_pair_min"].view(np.uint32))
        assert np.array_equal(pr["rank"], g[name + "_pair_imin"])
        nb = crd.shape[0]
        bonds = np.arange(nb - 1, dtype=np.int32)
        neg = np.full(nb - 1, -1, np.int32)
        per_item = np.sort(np.random.default_rng(2).uniform(0, 500, (nb - 1, n)), axis=1).astype(np.float32)
        po = eng.rank_match(np.stack([bonds, neg], 1), np.stack([bonds + 1, neg], 1), "min", per_item)
        assert np.array_equal(po["value"].view(np.uint32), g[name + "_poly_dist"].view(np.uint32))
        assert np.array_equal(po["rank"], g[name + "_poly_idx"])
        assert np.array_equal(po["matched"], np.take_along_axis(per_item, g[name + "_poly_idx"].astype(np.int64), 1))


@pytest.mark.parametrize("nstruct", [1, 2, 3, 127, 128, 129, 1000, 1025, 4097])
def test_rank_oracle_sizes_all_combinations_and_ties(nstruct):
    """Diploid pairs (all four combinations, as get_pair_dists' docstring intends), both
    reductions, ragged sizes; a third of the structures are exact duplicates, so ties are
    everywhere and must be ranked by structure index."""
    from igm_b200 import synthetic
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, nstruct, seed=nstruct, genome_scale=0.02)
    crd = pop.coordinates.copy()
    m = crd[:, 2::3].shape[1]
    crd[:, 2::3] = crd[:, 0::3][:, :m]            # structures 2, 5, 8, ... duplicate 0, 3, 6, ...
    pop = _pop(crd, pop.radii, pop.chrom, pop.copy_index.ptr, pop.copy_index.beads)
    ci = pop.copy_index
    rng = np.random.default_rng(7)
    loci = rng.integers(0, pop.n_hap, (24, 2))
    a = np.array([[ci[i][0], ci[i][1] if len(ci[i]) > 1 else -1] for i in loci[:, 0]], np.int32)
    b = np.array([[ci[j][0], ci[j][1] if len(ci[j]) > 1 else -1] for j in loci[:, 1]], np.int32)
    tgt = np.sort(rng.uniform(0, 3000, nstruct)).astype(np.float32)
    with ActdistEngine(pop, 0) as eng:
        for red in ("min", "max"):
            out = eng.rank_match(a, b, red, tgt)
            rad = eng.rank_match(a, None, red)
            for k, (i, j) in enumerate(loci):
                mind, maxd, imin, imax = ro.min_max_and_idx(ro.pair_values(crd, ci[i], ci[j]))
                v, r = (mind, imin) if red == "min" else (maxd, imax)
                assert np.array_equal(out["value"][k], v.astype(np.float32))
                assert np.array_equal(out["rank"][k], r)
                assert np.array_equal(out["matched"][k], tgt[r])
                mind, maxd, imin, imax = ro.min_max_and_idx(ro.radial_values(crd, ci[i]))
                v, r = (mind, imin) if red == "min" else (maxd, imax)
                assert np.array_equal(rad["value"][k], v.astype(np.float32)) and np.array_equal(rad["rank"][k], r)


def test_rank_match_rejects_bad_input():
    from igm_b200 import synthetic
    from igm_b200._lib import IgmkError
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, 16, seed=1, genome_scale=0.02)
    with ActdistEngine(pop, 0) as eng:
        with pytest.raises(IgmkError):
            eng.rank_match([[pop.nbead, -1]])
        with pytest.raises(IgmkError):
            eng.rank_match([[0, -1]], [[-1, -1]])
        with pytest.raises(ValueError):
            eng.rank_match([[0, -1]], None, "min", np.zeros(3, np.float32))
        assert eng.rank_match(np.zeros((0, 2), np.int32))["rank"].shape == (0, 16)


def _cfg(tmp_path, hss, extra_restraints, runtime):
    from igm_b200.steps._compat import Config
    return Config({"parameters": {"workdir": str(tmp_path), "tmp_dir": str(tmp_path / "tmp")},
                   "optimization": {"structure_output": hss},
                   "restraints": extra_restraints, "runtime": dict(runtime, opt_iter=2)})


def test_fish_step_end_to_end(tmp_path):
    from igm_b200 import hdf5, synthetic
    from igm_b200.steps import FishAssignmentStep
    pop = synthetic.make_population(2_000_000, 60, seed=21, genome_scale=0.02)
    hss = str(tmp_path / "pop.hss")
    pop.save_hss(hss)
    ci, n = pop.copy_index, pop.nstruct
    rng = np.random.default_rng(4)
    probes = rng.choice(pop.n_hap, 11, replace=False).astype(np.int32)
    pairs = rng.integers(0, pop.n_hap, (9, 2)).astype(np.int32)
    srt = lambda m, hi: np.sort(rng.uniform(0, hi, (m, n)), axis=1).astype(np.float32)
    inp = {"probes": probes, "pairs": pairs, "radial_min": srt(11, 4000), "radial_max": srt(11, 5000),
           "pair_min": srt(9, 3000), "pair_max": srt(9, 6000)}
    fish_in = str(tmp_path / "fish_in.h5")
    hdf5.write_h5(fish_in, inp)
    cfg = _cfg(tmp_path, hss, {"FISH": {"input_fish": fish_in, "tol_list": [200.0, 50.0], "batch_size": 4}}, {"FISH": {}})
    step = FishAssignmentStep(cfg)
    assert step.name() == "FishAssignmentStep (tol=200.00, iter=2)"
    step.run()
    out = cfg["runtime"]["FISH"]["fish_assignment_file"]
    assert out == str(tmp_path / "tmp" / "fish_actdist" / "fish_assignment.h5")
    with hdf5.open_h5(out) as f:
        got = {k: np.asarray(f[k][()]) for k in f.keys()}
    assert sorted(got) == sorted(inp)
    assert np.array_equal(got["probes"], probes) and np.array_equal(got["pairs"], pairs)
    for k, p in enumerate(probes):
        mind, maxd, imin, imax = ro.min_max_and_idx(ro.radial_values(pop.coordinates, ci[int(p)]))
        assert np.array_equal(got["radial_min"][k], inp["radial_min"][k][imin])
        assert np.array_equal(got["radial_max"][k], inp["radial_max"][k][imax])
    for k, (i, j) in enumerate(pairs):
        mind, maxd, imin, imax = ro.min_max_and_idx(ro.pair_values(pop.coordinates, ci[int(i)], ci[int(j)]))
        assert np.array_equal(got["pair_min"][k], inp["pair_min"][k][imin])
        assert np.array_equal(got["pair_max"][k], inp["pair_max"][k][imax])
    assert got["pair_min"].dtype == np.float32 and got["pairs"].dtype == np.int32
    # second iteration: the previous file is moved to the swap name (:334-338)
    cfg["runtime"]["FISH"].pop("tol")
    FishAssignmentStep(cfg).run()
    assert os.path.exists(out + ".tol_50.0000.iter_2") and os.path.exists(out)


def test_polymer_step_end_to_end(tmp_path):
    from igm_b200 import hdf5, synthetic
    from igm_b200.steps import PolymerAssignmentStep
    pop = synthetic.make_population(2_000_000, 50, seed=22, genome_scale=0.02)
    hss = str(tmp_path / "pop.hss")
    pop.save_hss(hss)
    edges = np.linspace(50.0, 900.0, 35)
    prob = np.random.default_rng(1).uniform(0.1, 1.0, 35)
    prob /= prob.sum()
    pfile = str(tmp_path / "poly.h5")
    hdf5.write_h5(pfile, {"bin_edges": edges, "probability": prob})
    cfg = _cfg(tmp_path, hss, {"polymer": {"polymer_file": pfile, "assignment_file": "polymer_assignment.h5"}},
               {"polymer": {}})
    step = PolymerAssignmentStep(cfg)
    assert step.name() == "PolymerAssignmentStep (iter=2)"
    np.random.seed(1234)
    step.run()
    out = cfg["runtime"]["polymer"]["assignment_file"]
    assert out == str(tmp_path / "tmp" / "poly_actdist" / "polymer_assignment.h5")
    with hdf5.open_h5(out) as f:
        loci, nn = np.asarray(f["loci"][()]), np.asarray(f["nn_dist"][()])
    assert np.array_equal(loci, np.arange(pop.nbead - 1)) and nn.shape == (pop.nbead - 1, pop.nstruct)
    np.random.seed(1234)                     # the reference's draws, PolymerAssignmentStep.py:115
    for i in range(pop.nbead - 1):
        sampled = np.sort(np.random.choice(edges, pop.nstruct, p=prob))
        _, idx = ro.polymer_dists(pop.coordinates, i)
        assert np.array_equal(nn[i], sampled[idx].astype(np.float32))
