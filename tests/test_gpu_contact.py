"""GPU parity tests of K2 (contact-frequency counts) against the NumPy oracle."""
import numpy as np
import pytest

from oracle import contact_oracle as co

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nstruct", [1, 31, 32, 100, 257])
@pytest.mark.parametrize("strict", [False, True])
def test_contact_counts_match_oracle(nstruct, strict):
    from igm_b200 import synthetic
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, nstruct, seed=50 + nstruct, genome_scale=0.012)
    nb = pop.nbead
    with ActdistEngine(pop, 0) as eng:
        full = eng.contact_counts(0, nb, 0, nb, 2.0, strict)
        exp = co.contact_counts_fast(pop.coordinates, pop.radii, np.arange(nb), np.arange(nb), 2.0, strict)
        assert np.array_equal(full, exp)
        assert np.array_equal(full, full.T)
        assert np.all(np.diag(full) == nstruct)      # d2 = 0 < rc^2 in both variants
        # ragged tile in the middle
        tile = eng.contact_counts(5, 37, 11, 45, 3.0, strict)
        exp = co.contact_counts_fast(pop.coordinates, pop.radii, np.arange(5, 42), np.arange(11, 56), 3.0, strict)
        assert np.array_equal(tile, exp)


def test_contact_counts_slow_oracle_spot():
    from igm_b200 import synthetic
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, 64, seed=9, genome_scale=0.008)
    with ActdistEngine(pop, 0) as eng:
        got = eng.contact_counts(0, 12, 3, 9, 2.0, False)
    exp = co.contact_counts(pop.coordinates, pop.radii, range(12), range(3, 12), 2.0, False)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("nstruct", [7, 100, 130])
def test_haploid_counts_are_copy_sums(nstruct):
    """Haploid mode == sumCopies projection of the bead-level map (oracle sum_copies),
    including haploid X/Y loci (one copy) and ragged tiles."""
    from igm_b200 import synthetic
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, nstruct, seed=70 + nstruct, genome_scale=0.012)
    nb, nh = pop.nbead, pop.n_hap
    full = co.contact_counts_fast(pop.coordinates, pop.radii, np.arange(nb), np.arange(nb), 2.0, False)
    exp = co.sum_copies(full, pop.copy_index.ptr, pop.copy_index.beads).astype(np.uint32)
    with ActdistEngine(pop, 0) as eng:
        got = eng.contact_counts_haploid(0, nh, 0, nh, 2.0, False)
        assert np.array_equal(got, exp)
        r1, c0 = min(nh, 43), min(nh - 1, 17)
        tile = eng.contact_counts_haploid(3, r1 - 3, c0, nh - c0, 2.0, False)
        assert np.array_equal(tile, exp[3:r1, c0:])
        from igm_b200.contact import haploid_contact_counts
        dense = haploid_contact_counts(eng, 2.0, False, block=37)
        assert np.array_equal(dense, exp)
        fulls = co.contact_counts_fast(pop.coordinates, pop.radii, np.arange(nb), np.arange(nb), 3.0, True)
        exps = co.sum_copies(fulls, pop.copy_index.ptr, pop.copy_index.beads).astype(np.uint32)
        assert np.array_equal(eng.contact_counts_haploid(0, nh, 0, nh, 3.0, True), exps)      # strict '<'
        # two "ranks" computed separately add up to the whole
        a = haploid_contact_counts(eng, 2.0, False, block=16, rank=0, world=2)
        b = haploid_contact_counts(eng, 2.0, False, block=16, rank=1, world=2)
        assert np.array_equal(np.maximum(a, b), exp)


def test_get_simulated_hic_and_evaluation_step(tmp_path):
    from igm_b200 import synthetic
    from igm_b200.contact import get_simulated_hic
    from igm_b200.population import ProbMatrix
    from igm_b200.steps._compat import Config
    from igm_b200.steps.HicEvaluationStep import HicEvaluationStep, eps
    pop = synthetic.make_population(2_000_000, 60, seed=4, genome_scale=0.012)
    hss = str(tmp_path / "pop.hss")
    pop.save_hss(hss)
    pm = get_simulated_hic(hss, 2.0)
    nb = pop.nbead
    full = co.contact_counts_fast(pop.coordinates, pop.radii, np.arange(nb), np.arange(nb), 2.0, False)
    exp = co.sum_copies(full, pop.copy_index.ptr, pop.copy_index.beads) / pop.nstruct
    dense = np.zeros((pm.n, pm.n))
    dense[pm.rows(), pm.indices] = pm.data
    assert np.array_equal(dense.astype(np.float32), np.triu(exp, 1).astype(np.float32))
    # evaluation step: out_matrix.hcs = clipped haploid map at cr * (1 + eps)
    inp = str(tmp_path / "in.hcs")
    pm.save_hcs(inp)
    cfg = Config({"parameters": {"workdir": str(tmp_path), "tmp_dir": "tmp"},
                  "optimization": {"structure_output": hss},
                  "restraints": {"Hi-C": {"input_matrix": inp, "contact_range": 2.0}},
                  "runtime": {"Hi-C": {"intra_sigma": 0.05, "inter_sigma": 0.05}, "opt_iter": 3}})
    step = HicEvaluationStep(cfg)
    assert step.name() == "HicEvaluationStep (sigma=5.00%, iter=3)"
    step.run()
    out = ProbMatrix.from_hcs(str(tmp_path / "evaluation" / "Hi-C" / "sigma_5.00.iter_3" / "out_matrix.hcs"))
    full2 = co.contact_counts_fast(pop.coordinates, pop.radii, np.arange(nb), np.arange(nb), 2.0 * (1 + eps), False)
    exp2 = (co.sum_copies(full2, pop.copy_index.ptr, pop.copy_index.beads) / pop.nstruct).clip(0, 1)
    dense2 = np.zeros((out.n, out.n))
    dense2[out.rows(), out.indices] = out.data
    assert np.array_equal(dense2.astype(np.float32), np.triu(exp2, 1).astype(np.float32))
    stats = open(str(tmp_path / "evaluation" / "Hi-C" / "sigma_5.00.iter_3" / "stats.txt")).read().splitlines()
    assert stats[0].startswith("#score") and len(stats[1].split()) == 3
