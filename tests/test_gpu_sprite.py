"""GPU parity of K4 (SPRITE Rg^2 with exhaustive copy choice, next row f3) against the
reference's own compiled get_rg2s_cpp (oracle/_ref/libsprite_ref.so) and its KATs."""
import numpy as np
import pytest

from oracle import sprite_oracle as so
from tests.test_sprite_cpu import KATS

pytestmark = pytest.mark.gpu


def _oracle(crd, copies):
    return so.ref_get_rgs2(crd, np.array(copies, np.int32)) if so.ref_available() else so.get_rgs2_port(crd, copies)


@pytest.mark.parametrize("case", range(3))
def test_known_answers(case):
    from igm_b200.engine import ActdistEngine
    crd, cn, (rg, best, ci) = KATS[case]
    with ActdistEngine(nbead=crd.shape[0], nstruct=crd.shape[1], device=0) as eng:
        eng.upload_coordinates(crd)
        regions, b = [], 0
        for k in cn:
            regions.append(list(range(b, b + k)))
            b += k
        (r, bs, c), = eng.sprite_rg2([regions])
    assert np.allclose(r, rg, rtol=1e-6) and bs == best and c.tolist() == ci


@pytest.mark.parametrize("nstruct", [1, 57, 1000])
def test_random_clusters_vs_reference(nstruct):
    from igm_b200 import synthetic
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, nstruct, seed=20 + nstruct, genome_scale=0.02)
    rng = np.random.default_rng(nstruct)
    ci = pop.copy_index
    clusters = []
    for _ in range(40):
        m = int(rng.integers(1, 7))
        loci = rng.choice(pop.n_hap, size=m, replace=False)
        clusters.append([ci[int(l)] for l in loci])            # copies of each locus
    clusters.append([[0, 1, 2], [5], [7, 9]])                   # three alternative locations
    with ActdistEngine(pop, 0) as eng:
        got = eng.sprite_rg2(clusters)
    for cl, (r, bs, c) in zip(clusters, got):
        beads = [b for reg in cl for b in reg]
        crd = np.ascontiguousarray(pop.coordinates[beads])      # (B, N, 3): what get_rgs2 receives
        er, ebs, ec = _oracle(crd, [len(reg) for reg in cl])
        assert np.array_equal(r.view(np.uint32), er.view(np.uint32))
        assert bs == ebs and np.array_equal(c, ec)


def test_limits_and_errors():
    from igm_b200 import _lib
    from igm_b200.engine import ActdistEngine
    crd = np.zeros((60, 4, 3), np.float32)
    with ActdistEngine(nbead=60, nstruct=4, device=0) as eng:
        eng.upload_coordinates(crd)
        with pytest.raises(_lib.IgmkError):
            eng.sprite_rg2([[[b] for b in range(30)]])          # more than 24 regions
        with pytest.raises(_lib.IgmkError):
            eng.sprite_rg2([[[0, 99]]])                         # bead id out of range
        (r, bs, c), = eng.sprite_rg2([[[0], [1]]])              # all points coincide: Rg^2 = 0
        assert np.all(r == 0) and bs == 0
