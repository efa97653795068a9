"""GPU parity of K3 (M-step Hi-C restraint selection, next row f2) through the C ABI."""
import numpy as np
import pytest

from oracle import restraint_oracle as ro
from tests.test_restraint_cpu import GOLDEN

pytestmark = pytest.mark.gpu


def _unpack(bitmap, nstruct):
    bits = np.unpackbits(bitmap.view(np.uint8), axis=1, bitorder="little")
    return bits[:, :nstruct].astype(bool)


@pytest.mark.parametrize("kind", ["intra", "inter"])
def test_restraint_golden(kind):
    from igm_b200.engine import ActdistEngine
    g = np.load(GOLDEN)
    coords = g["coords"]
    nb, ns = coords.shape[0], coords.shape[1]
    exp = np.unpackbits(g["sel_" + kind], axis=1)[:, :ns].astype(bool)
    with ActdistEngine(nbead=nb, nstruct=ns, device=0) as eng:
        eng.upload_coordinates(coords)
        eng.set_bead_chrom(g["chrom"])
        bitmap, counts = eng.restraint_select(g["row"], g["col"], g["dist"], kind)
        got = _unpack(bitmap, ns)
        assert np.array_equal(got, exp)
        assert np.array_equal(counts, exp.sum(axis=1))
        assert not bitmap.view(np.uint8)[:, (ns + 7) // 8 + 1:].any()      # padding bits stay clear
        s = 13
        assert np.array_equal(eng.records_of_structure(bitmap, s), np.nonzero(exp[:, s])[0])


@pytest.mark.parametrize("nstruct", [1, 129, 1000])
def test_restraint_oracle_sizes(nstruct):
    from igm_b200 import synthetic
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, nstruct, seed=3 + nstruct, genome_scale=0.01)
    rng = np.random.default_rng(nstruct)
    n = 150
    row = rng.integers(0, pop.nbead, n).astype(np.int32)
    col = rng.integers(0, pop.nbead, n).astype(np.int32)
    sp = rng.integers(0, nstruct, n)
    d_true = np.array([np.linalg.norm(pop.coordinates[a, s] - pop.coordinates[b, s]) for a, b, s in zip(row, col, sp)])
    dist = d_true.astype(np.float32)                    # exact ties in at least one structure
    dist[1::2] = (d_true[1::2] * rng.uniform(0.5, 1.5, len(d_true[1::2]))).astype(np.float32)
    dist[5] = 0.0
    dist[6] = np.inf
    with ActdistEngine(pop, 0) as eng:
        for kind in ("intra", "inter", "any"):
            bitmap, counts = eng.restraint_select(row, col, dist, kind)
            exp = ro.select_bitmap(pop.coordinates, pop.chrom, row, col, dist, kind)
            assert np.array_equal(_unpack(bitmap, nstruct), exp), kind
            assert np.array_equal(counts, exp.sum(axis=1))
