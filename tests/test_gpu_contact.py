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
