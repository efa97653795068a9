"""GPU parity of K4 (SPRITE Rg^2 with exhaustive copy choice, next row f3) against the
reference's own compiled get_rg2s_cpp (oracle/_ref/libsprite_ref.so) and its KATs."""
import os

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


# ------------------------------------------------ cluster-level driver and the Step drop-in
def _engine_for(crd, chrom, copy_ptr_beads):
    from igm_b200.engine import ActdistEngine
    from igm_b200.population import CopyIndex
    ptr, beads = copy_ptr_beads
    eng = ActdistEngine(nbead=crd.shape[0], nstruct=crd.shape[1], device=0)
    eng.upload_coordinates(crd)
    ci = CopyIndex(ptr, beads)
    eng.set_index(ci.ptr, ci.beads, np.ascontiguousarray(chrom[:len(ci)], np.int32),
                  np.ones(crd.shape[0], np.float32))
    return eng, ci


def test_batch_gyration_radii_equals_reference_golden():
    """compute_gyration_radius on the GPU (K4 + K4b) against what the reference's own
    compiled Cython function returned, cluster by cluster under the recorded seeds, and
    as one batch against the pinned oracle port under one shared seed."""
    from igm_b200.steps.SpriteAssignmentStep import batch_gyration_radii
    from tests import helpers as H
    for name, crd, chrom, cidict, clusters, rg2s, best, sel, seed0, raw in H.sprite_golden_cases():
        eng, ci = _engine_for(crd, chrom, raw)
        with eng:
            allind = np.arange(crd.shape[1])
            for k, cl in enumerate(clusters):
                np.random.seed(seed0 + k)
                (rg, select), = batch_gyration_radii(eng, chrom, ci, [cl], 99)
                assert np.array_equal(np.asarray(rg, np.float32).view(np.uint32), rg2s[k].view(np.uint32)), (name, k)
                assert np.array_equal(select(allind), sel[k]), (name, k)
            # one batch, shared random stream; clusters over 2 chromosomes are skipped
            np.random.seed(7)
            exp = []
            for cl in clusters:
                if len(np.unique(chrom[cl])) > 2:
                    exp.append(None)
                else:
                    exp.append(so.compute_gyration_radius_port(crd, cl, chrom, cidict))
            np.random.seed(7)
            got = batch_gyration_radii(eng, chrom, ci, clusters, 2)
            assert sum(e is None for e in exp) > 0 and sum(e is not None for e in exp) > 3
            for e, g in zip(exp, got):
                assert (e is None) == (g is None)
                if e is not None:
                    assert np.array_equal(np.asarray(e[0], np.float32).view(np.uint32),
                                          np.asarray(g[0], np.float32).view(np.uint32))
                    assert np.array_equal(g[1](allind), e[2])


def test_sprite_step_end_to_end(tmp_path):
    """SpriteAssignmentStep (setup -> task -> reduce) on files, against the restated reference
    flow driven by the pinned oracle, same NumPy seed: identical assignment.h5."""
    from igm_b200 import hdf5, synthetic
    from igm_b200.steps import SpriteAssignmentStep
    from igm_b200.steps._compat import Config
    from tests import helpers as H
    pop = synthetic.make_population(2_000_000, 130, seed=17, genome_scale=0.03)
    hss = str(tmp_path / "igm-model.hss")
    pop.save_hss(hss)
    rng = np.random.default_rng(2)
    chrom_hap = pop.chrom_hap()
    n_hap = pop.n_hap
    clusters = []
    for k in range(23):
        if k % 4 == 0:
            pool = np.nonzero(chrom_hap == rng.choice(np.unique(chrom_hap)))[0]
        elif k % 4 == 3:
            pool = np.arange(n_hap)                     # many chromosomes: skipped by max_chrom
        else:
            cs = rng.choice(np.unique(chrom_hap), size=int(rng.integers(2, 5)), replace=False)
            pool = np.nonzero(np.isin(chrom_hap, cs))[0]
        size = int(min(len(pool), rng.integers(2, 30)))
        clusters.append(np.sort(rng.choice(pool, size=size, replace=False)).astype(np.int32))
    indptr = np.concatenate([[0], np.cumsum([len(c) for c in clusters])]).astype(np.int32)
    clf = str(tmp_path / "clusters.h5")
    hdf5.write_h5(clf, {"indptr": indptr, "data": np.concatenate(clusters).astype(np.int32)})
    cfg = Config({"parameters": {"workdir": str(tmp_path), "tmp_dir": str(tmp_path / "tmp")},
                  "optimization": {"structure_output": hss},
                  "restraints": {"sprite": {"clusters": clf, "volume_fraction_list": [0.05], "batch_size": 5,
                                            "keep_best": 20, "max_chrom_in_cluster": 4, "radius_kt": 80.0}},
                  "runtime": {"sprite": {}, "opt_iter": 3}})
    step = SpriteAssignmentStep(cfg)
    assert step.name() == "SpriteAssignmentStep (volume_fraction=0.1%, iter=3)"
    np.random.seed(2024)
    step.setup()
    assert list(step.argument_list) == [0, 1, 2, 3, 4] and step.n_clusters == 23 and step.n_struct == 130
    for b in step.argument_list:
        step.task(b, cfg, step.tmp_dir)
    step.reduce()
    with hdf5.open_h5(os.path.join(step.tmp_dir, "assignment.h5")) as f:
        got_assign, got_sel = np.asarray(f["assignment"][()]), np.asarray(f["selected"][()])
        assert f["assignment"].dtype == np.int32 and f["selected"].dtype == np.int32
        assert np.array_equal(np.asarray(f["indptr"][()]), indptr)
    # the reference flow on the CPU, same random stream
    np.random.seed(2024)
    chrom_bead = pop.chrom
    cidict = {i: pop.copy_index[i] for i in range(n_hap)}
    batches = [H.sprite_reference_task(pop.coordinates, chrom_bead, cidict, clusters[b * 5:(b + 1) * 5], 20, 4)
               for b in range(5)]
    exp_assign, exp_sel = H.sprite_reference_reduce(batches, indptr, 130, 5, 80.0)
    assert np.array_equal(got_assign, exp_assign)
    assert np.array_equal(got_sel, exp_sel)
    assert (got_assign >= 0).sum() >= 12 and (got_assign == -1).sum() >= 3
