"""GPU tests of igmk_actdist_host_population: the population upload pipelined with the pair
kernels (slices of the list from its end, beads uploaded as the slices need them) gives the
bytes of igmk_upload_coords + igmk_actdist_host - for sorted and shuffled lists, for the
usual copy index (copy 0 block, then copy 1 block) and for a locus-major one, for
populations on both sides of the 1024-structure boundary - and leaves the context fully
staged.  Checked against the oracle too."""
import numpy as np
import pytest

from oracle import actdist_oracle as orc
from tests import helpers as H
from tests.test_gpu_actdist import _check_against_details, _sorted_pairs

pytestmark = pytest.mark.gpu


def _fresh_engine(pop):
    """Index set, NO coordinates (the context's HBM copy is all zeros)."""
    from igm_b200.engine import ActdistEngine
    eng = ActdistEngine(nbead=pop.nbead, nstruct=pop.nstruct, device=0)
    eng.set_index(pop.copy_index.ptr, pop.copy_index.beads, pop.chrom_hap(), pop.radii)
    eng.set_bead_chrom(pop.chrom)
    return eng


def _inputs(pop, n, seed):
    rng = np.random.default_rng(seed)
    ii, jj = _sorted_pairs(rng, pop.n_hap, n)
    pw = np.exp(rng.uniform(np.log(0.004), np.log(0.3), len(ii))).astype(np.float32).astype(np.float64)
    pl = np.where(rng.random(len(ii)) < 0.5, 0.0, orc.text_roundtrip(rng.uniform(0, 0.05, len(ii))).astype(np.float64))
    return rng, ii, jj, pw, pl


@pytest.mark.parametrize("nstruct,npairs,slice_pairs", [(300, 60000, 4096), (1000, 90000, 8192), (2600, 6000, 1024)])
@pytest.mark.parametrize("order", ["sorted", "shuffled", "descending"])
def test_pipelined_population_matches_two_calls(nstruct, npairs, slice_pairs, order, monkeypatch):
    from igm_b200 import synthetic
    from igm_b200.engine import ActdistEngine, pinned_array
    monkeypatch.setenv("IGMK_HOST_SLICE", str(slice_pairs))
    pop = synthetic.make_population(2_000_000, nstruct, seed=900 + nstruct, genome_scale=0.05)
    rng, ii, jj, pw, pl = _inputs(pop, npairs, nstruct)
    if order == "shuffled":
        perm = rng.permutation(len(ii))
        ii, jj, pw, pl = ii[perm], jj[perm], pw[perm], pl[perm]
    elif order == "descending":
        ii, jj, pw, pl = ii[::-1].copy(), jj[::-1].copy(), pw[::-1].copy(), pl[::-1].copy()
    xyz = pinned_array(pop.coordinates.shape, np.float32, tag="test_pipeline_xyz")
    xyz[...] = pop.coordinates
    with ActdistEngine(pop, device=0) as eng:
        want = eng.actdist(ii, jj, pw, pl, 2.0, 1, "lb", 0)
    with _fresh_engine(pop) as eng:
        got = eng.actdist_with_population(xyz, ii, jj, pw, pl, 2.0, 1, "lb", 0)
        assert got.tobytes() == want.tobytes()
        # the context is fully staged afterwards: a plain call and a contact tile agree
        again = eng.actdist(ii, jj, pw, pl, 2.0, 1, "lb", 0)
        assert again.tobytes() == want.tobytes()
        nb = min(pop.nbead, 96)
        tile = eng.contact_counts(pop.nbead - nb, nb, 0, nb)
    with ActdistEngine(pop, device=0) as eng:
        assert np.array_equal(tile, eng.contact_counts(pop.nbead - nb, nb, 0, nb))
    sel = np.sort(rng.choice(len(ii), 120, replace=False))
    _, dets = orc.run_pairs(ii[sel], jj[sel], pw[sel], pl[sel], pop.coordinates, pop.radii, pop.chrom_hap(),
                            pop.copy_index, 1, 2.0, orc.MODE_LB)
    _check_against_details(got[sel], dets)


def test_pipelined_population_locus_major_index(monkeypatch):
    """Beads stored locus by locus (copy 0, copy 1, copy 0, ...): the two copy regions
    interleave and the upload falls back to one descending cursor."""
    from igm_b200 import synthetic
    from igm_b200.engine import ActdistEngine
    from igm_b200.population import CopyIndex, Population
    monkeypatch.setenv("IGMK_HOST_SLICE", "2048")
    base = synthetic.make_population(2_000_000, 500, seed=77, genome_scale=0.05)
    # new bead order: for every locus its copies, consecutively
    ptr, beads = np.asarray(base.copy_index.ptr), np.asarray(base.copy_index.beads)
    order = beads.copy()                                   # new position k holds old bead order[k]
    new_of_old = np.empty(base.nbead, np.int64)
    new_of_old[order] = np.arange(len(order))
    assert len(order) == base.nbead
    # the reference reads the chromosome of locus l at index.chrom[l] (ActivationDistanceStep.py:393):
    # keep that table as it was, whatever bead now sits at position l
    chrom = base.chrom[order].copy()
    chrom[:base.n_hap] = base.chrom_hap()
    pop = Population(np.ascontiguousarray(base.coordinates[order]), base.radii[order], chrom,
                     CopyIndex(ptr, new_of_old[beads].astype(np.int32)))
    rng, ii, jj, pw, pl = _inputs(pop, 30000, 5)
    with ActdistEngine(pop, device=0) as eng:
        want = eng.actdist(ii, jj, pw, pl, 2.0, 0, "gp", 0)
    with _fresh_engine(pop) as eng:
        got = eng.actdist_with_population(pop.coordinates, ii, jj, pw, pl, 2.0, 0, "gp", 0)
    assert got.tobytes() == want.tobytes()
    sel = np.sort(rng.choice(len(ii), 80, replace=False))
    _, dets = orc.run_pairs(ii[sel], jj[sel], pw[sel], pl[sel], pop.coordinates, pop.radii, pop.chrom_hap(),
                            pop.copy_index, 0, 2.0, orc.MODE_GP)
    _check_against_details(got[sel], dets)


def test_pipelined_population_edge_cases():
    from igm_b200 import synthetic
    from igm_b200._lib import IgmkError
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(5_000_000, 130, seed=3, genome_scale=0.05)
    rng, ii, jj, pw, pl = _inputs(pop, 500, 1)
    with ActdistEngine(pop, device=0) as eng:
        want = eng.actdist(ii, jj, pw, pl)
    # an empty list just stages the population
    with _fresh_engine(pop) as eng:
        out = eng.actdist_with_population(pop.coordinates, ii[:0], jj[:0], pw[:0], pl[:0])
        assert len(out) == 0
        assert eng.actdist(ii, jj, pw, pl).tobytes() == want.tobytes()
    # one pair touching only the last locus, then the whole list on the same context
    with _fresh_engine(pop) as eng:
        last = np.array([pop.n_hap - 2], np.int32), np.array([pop.n_hap - 1], np.int32)
        one = eng.actdist_with_population(pop.coordinates, last[0], last[1], np.array([0.3]))
        with ActdistEngine(pop, device=0) as ref:
            assert one.tobytes() == ref.actdist(last[0], last[1], np.array([0.3])).tobytes()
        assert eng.actdist(ii, jj, pw, pl).tobytes() == want.tobytes()
    # no index: refused
    eng = ActdistEngine(nbead=pop.nbead, nstruct=pop.nstruct, device=0)
    with eng:
        eng.n_hap = pop.n_hap
        eng._ncopies = np.diff(pop.copy_index.ptr)
        eng._chrom_hap = pop.chrom_hap()
        eng._uniform_ploidy = True
        with pytest.raises(IgmkError):
            eng.actdist_with_population(pop.coordinates, ii, jj, pw, pl)


def test_population_shares_between_contexts():
    """NVLink replication path on whatever devices the box has: engine B uploads half of the
    beads and takes the other half from engine A (igmk_copy_coords_peer); the staged rows
    (igmk_coords_device, zero-copy torch view) and the A-step results are identical.  With
    >= 2 GPUs the two engines sit on different devices."""
    import torch
    from igm_b200 import synthetic
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, 200, seed=21, genome_scale=0.05)
    rng, ii, jj, pw, pl = _inputs(pop, 4000, 2)
    dev_b = 1 if torch.cuda.device_count() > 1 else 0
    half = pop.nbead // 2
    with ActdistEngine(pop, device=0) as a:
        with ActdistEngine(nbead=pop.nbead, nstruct=pop.nstruct, device=dev_b) as b:
            b.set_index(pop.copy_index.ptr, pop.copy_index.beads, pop.chrom_hap(), pop.radii)
            b.upload_coordinates(pop.coordinates[:half], bead0=0)
            b.copy_coordinates_from(a, half, pop.nbead - half)
            ta, tb = a.coords_tensor(), b.coords_tensor()
            assert ta.shape == (pop.nbead + 1 + 64, 3 * 256) and tb.shape == ta.shape
            assert torch.equal(ta.cpu(), tb.cpu())
            assert float(ta[pop.nbead:].abs().max()) == 0.0          # origin row + spare rows
            assert b.actdist(ii, jj, pw, pl).tobytes() == a.actdist(ii, jj, pw, pl).tobytes()
            from igm_b200._lib import IgmkError
            with pytest.raises(IgmkError):
                b.copy_coordinates_from(a, pop.nbead - 1, 2)
