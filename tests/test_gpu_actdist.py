"""GPU parity tests of K1 (activation distance) through the C ABI.

The CUDA path is compared with (a) the committed golden vectors produced by
the reference's own get_actdist and (b) the NumPy oracle on seeded inputs.
Bar: bit-exact d2 / count / o / p, identical record text (sha256).
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import actdist_oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu

MODES = {"lb": orc.MODE_LB, "gp": orc.MODE_GP}


def _engine(pop):
    from igm_b200.engine import ActdistEngine
    return ActdistEngine(pop, device=0)


def _records_text(eng, ii, jj, res):
    row, col, dist, prob = eng.expand_records(ii, jj, res)
    return row, col, dist, prob


def _check_against_details(res, dets):
    exp = orc.details_to_arrays(dets)
    assert np.array_equal(res["contact_count"], exp["contact_count"])
    assert np.array_equal(res["o"], exp["o"])
    assert np.array_equal(res["p"].view(np.uint64), exp["p"].view(np.uint64))
    sel = exp["o"] >= 0
    assert np.array_equal(res["d2_sel_bits"][sel], exp["d2_sel_bits"][sel])
    assert np.array_equal(res["nrec"], exp["nrec"])


@pytest.mark.parametrize("case", [c[0] for c in H.all_small_cases()])
@pytest.mark.parametrize("algo", [0, 1])
def test_golden_small(case, algo):
    name, npz, prefix = [c for c in H.all_small_cases() if c[0] == case][0]
    pop, ii, jj, pw, pl = H.load_case(npz, prefix)
    with _engine(pop) as eng:
        for mode in ("lb", "gp"):
            for it_corr in (0, 1):
                g = H.golden_out(npz, prefix, mode, it_corr)
                res = eng.actdist(ii, jj, pw, pl, 2.0, it_corr, mode, algo)
                assert np.array_equal(res["nrec"], g["nrec"]), (mode, it_corr)
                row, col, dist, prob = eng.expand_records(ii, jj, res)
                assert np.array_equal(row, g["row"]) and np.array_equal(col, g["col"])
                # activation distance: float64 sqrt of the selected float32 d2 (bit-exact)
                first = np.concatenate([[0], np.cumsum(g["nrec"])[:-1]])[g["nrec"] > 0]
                ad = np.sqrt(res["d2_sel_bits"][g["nrec"] > 0].view(np.float32).astype(np.float64))
                assert np.array_equal(ad.view(np.uint64), g["ad"][first].view(np.uint64))
                assert np.array_equal(res["p"][g["nrec"] > 0].view(np.uint64), g["p"][first].view(np.uint64))
                # stored columns = 4-decimal text round trip of the reference output
                assert np.array_equal(dist.view(np.uint32), orc.text_roundtrip(g["ad"]).view(np.uint32))
                assert np.array_equal(prob.view(np.uint32), orc.text_roundtrip(g["p"]).view(np.uint32))
                # and the '%d.out.tmp' text itself
                recs = list(zip(row.tolist(), col.tolist(), np.repeat(ad, g["nrec"][g["nrec"] > 0]).tolist(),
                                np.repeat(res["p"][g["nrec"] > 0], g["nrec"][g["nrec"] > 0]).tolist()))
                assert hashlib.sha256(orc.task_text(recs).encode()).hexdigest() == g["sha"]


def test_contact_range_variant():
    s = np.load(os.path.join(H.GOLDEN, "synth_small.npz"))
    pop, ii, jj, pw, pl = H.load_case(s, "n100")
    with _engine(pop) as eng:
        for mode in ("lb", "gp"):
            g = H.golden_out(s, "n100cr35", mode, 1)
            res = eng.actdist(ii, jj, pw, pl, 3.5, 1, mode)
            row, col, dist, prob = eng.expand_records(ii, jj, res)
            assert np.array_equal(dist.view(np.uint32), orc.text_roundtrip(g["ad"]).view(np.uint32))
            assert np.array_equal(prob.view(np.uint32), orc.text_roundtrip(g["p"]).view(np.uint32))


@pytest.mark.parametrize("nstruct", [5, 64, 129, 500, 1000, 1024])
@pytest.mark.parametrize("mode", ["lb", "gp"])
def test_oracle_synthetic_warp_mode(nstruct, mode):
    from igm_b200 import synthetic
    pop = synthetic.make_population(2_000_000, nstruct, seed=100 + nstruct, genome_scale=0.02)
    rng = np.random.default_rng(nstruct)
    nh = pop.n_hap
    ii = rng.integers(0, nh, 300)
    jj = rng.integers(0, nh, 300)
    k = ii < jj
    ii, jj = ii[k].astype(np.int32), jj[k].astype(np.int32)
    pw = rng.uniform(0.001, 1.0, len(ii)).astype(np.float32).astype(np.float64)
    pw[::9] = 1.0
    pl = np.where(rng.random(len(ii)) < 0.5, 0.0, orc.text_roundtrip(rng.uniform(0, 0.8, len(ii))).astype(np.float64))
    with _engine(pop) as eng:
        for it_corr in (0, 1):
            _, dets = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                    pop.copy_index, it_corr, 2.0, MODES[mode])
            for algo in (0, 1):
                res = eng.actdist(ii, jj, pw, pl, 2.0, it_corr, mode, algo)
                _check_against_details(res, dets)


@pytest.mark.parametrize("nstruct,group_threads", [(1000, 64), (1500, 0), (2000, 32), (2600, 0), (4100, 32), (4100, 64),
                                                   (4100, 160), (4100, 512), (10000, 0), (12288, 0), (20000, 0)])
def test_oracle_synthetic_block_mode(nstruct, group_threads, monkeypatch):
    """CTA-per-pair groups (and forced group sizes, incl. one warp for large N)."""
    from igm_b200 import synthetic
    if group_threads:
        monkeypatch.setenv("IGMK_GROUP_THREADS", str(group_threads))
    pop = synthetic.make_population(2_000_000, nstruct, seed=7 + nstruct, genome_scale=0.004)
    rng = np.random.default_rng(nstruct)
    nh = pop.n_hap
    ii = rng.integers(0, nh, 120)
    jj = rng.integers(0, nh, 120)
    k = ii < jj
    ii, jj = ii[k].astype(np.int32), jj[k].astype(np.int32)
    pw = rng.uniform(0.001, 1.0, len(ii)).astype(np.float32).astype(np.float64)
    pl = np.zeros(len(ii))
    with _engine(pop) as eng:
        for mode in ("lb", "gp"):
            _, dets = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                    pop.copy_index, 1, 2.0, MODES[mode])
            res = eng.actdist(ii, jj, pw, pl, 2.0, 1, mode, 0)
            _check_against_details(res, dets)


@pytest.mark.parametrize("nstruct,tile_block", [(300, 0), (300, 7), (1000, 64), (2600, 0)])
def test_jblock_order_and_locus_tile(nstruct, tile_block, monkeypatch):
    """Forced J-block processing order (tiny L2 budget -> many blocks) and the
    shared-memory locus-i tile with short CTA blocks; results stay in input order."""
    from igm_b200 import synthetic
    monkeypatch.setenv("IGMK_L2_BUDGET", str(24 * 1024 * 7))
    monkeypatch.setenv("IGMK_ORDER_MIN", "1")
    monkeypatch.setenv("IGMK_TILE_BLOCK", str(tile_block))
    pop = synthetic.make_population(2_000_000, nstruct, seed=31 + nstruct, genome_scale=0.03)
    rng = np.random.default_rng(nstruct + tile_block)
    nh = pop.n_hap
    # CSR-like list: runs of equal i with many j, as setup() produces
    ii = np.repeat(np.arange(0, nh - 1, 2), 9)
    jj = (ii + 1 + rng.integers(0, nh, len(ii))) % nh
    k = ii != jj
    ii, jj = ii[k].astype(np.int32), jj[k].astype(np.int32)
    lo, hi = np.minimum(ii, jj), np.maximum(ii, jj)
    nc = pop.copy_index.ncopies()
    ch = pop.chrom_hap()
    ok = ~((ch[lo] == ch[hi]) & (nc[lo] != nc[hi]))
    ii, jj = lo[ok], hi[ok]
    pw = rng.uniform(0.001, 1.0, len(ii)).astype(np.float32).astype(np.float64)
    pl = np.zeros(len(ii))
    with _engine(pop) as eng:
        for mode in ("lb", "gp"):
            _, dets = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                    pop.copy_index, 1, 2.0, MODES[mode])
            res = eng.actdist(ii, jj, pw, pl, 2.0, 1, mode, 0)
            _check_against_details(res, dets)


@pytest.mark.parametrize("block_stop", [32, 100, 512, 1024])
@pytest.mark.parametrize("nstruct", [1500, 10000])
def test_block_stop_variants(nstruct, block_stop, monkeypatch):
    """CTA groups: the key bisection stops at <= IGMK_BLOCK_STOP candidates, which are then
    compacted without atomics and bisected on their full float32 patterns."""
    from igm_b200 import synthetic
    monkeypatch.setenv("IGMK_BLOCK_STOP", str(block_stop))
    pop = synthetic.make_population(2_000_000, nstruct, seed=900 + nstruct, genome_scale=0.004)
    rng = np.random.default_rng(nstruct + block_stop)
    nh = pop.n_hap
    ii = rng.integers(0, nh, 150)
    jj = rng.integers(0, nh, 150)
    k = ii < jj
    ii, jj = ii[k].astype(np.int32), jj[k].astype(np.int32)
    pw = np.concatenate([rng.uniform(0.0005, 1.0, len(ii) - 3), [1.0, 0.9999, 1e-4]])
    pl = np.zeros(len(ii))
    with _engine(pop) as eng:
        for mode in ("lb", "gp"):
            _, dets = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                    pop.copy_index, 0, 2.0, MODES[mode])
            res = eng.actdist(ii, jj, pw, pl, 2.0, 0, mode, 0)
            _check_against_details(res, dets)


@pytest.mark.parametrize("nconf", [1, 3, 40])
def test_block_groups_many_equal_distances(nconf, monkeypatch):
    """CTA groups on populations whose structures are copies of a few conformations: every
    distance value occurs hundreds of times (fat keys, lists full of duplicates, the
    all-equal exits of both bisections)."""
    from igm_b200.population import Population
    from igm_b200 import synthetic
    nstruct = 2100
    bins = np.array([6, 5, 3])
    chrom_hap, chrom_bead, copy_bead, ci = synthetic.build_index(bins, n_diploid_chroms=2)
    nbead = len(chrom_bead)
    rng = np.random.default_rng(nconf)
    conf = (rng.standard_normal((nbead, nconf, 3)) * 300.0).astype(np.float32)
    crd = np.ascontiguousarray(conf[:, rng.integers(0, nconf, nstruct), :])
    pop = Population(crd, np.full(nbead, 50.0, np.float32), chrom_bead, ci, copy_bead)
    ii, jj = np.triu_indices(pop.n_hap, 1)
    ii, jj = ii.astype(np.int32), jj.astype(np.int32)
    pw = rng.uniform(0.001, 1.0, len(ii))
    pl = np.zeros(len(ii))
    for stop in ("64", "1024"):
        monkeypatch.setenv("IGMK_BLOCK_STOP", stop)
        with _engine(pop) as eng:
            for mode in ("lb", "gp"):
                _, dets = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                        pop.copy_index, 0, 2.0, MODES[mode])
                res = eng.actdist(ii, jj, pw, pl, 2.0, 0, mode, 0)
                _check_against_details(res, dets)


def _sorted_pairs(rng, nh, n):
    ii = rng.integers(0, nh, n)
    jj = rng.integers(0, nh, n)
    k = ii < jj
    ii, jj = ii[k].astype(np.int32), jj[k].astype(np.int32)
    order = np.lexsort((jj, ii))
    return ii[order], jj[order]


def test_dynamic_blocks_variant_matches(monkeypatch):
    """Key-array kernels alone (IGMK_LIST=0): pair blocks handed out by a device-wide counter
    instead of round-robin give byte-identical results."""
    from igm_b200 import synthetic
    pop = synthetic.make_population(2_000_000, 700, seed=41, genome_scale=0.05)
    rng = np.random.default_rng(8)
    ii, jj = _sorted_pairs(rng, pop.n_hap, 70000)
    pw = rng.uniform(0.001, 1.0, len(ii))
    monkeypatch.setenv("IGMK_LIST", "0")
    with _engine(pop) as eng:
        base = eng.actdist(ii, jj, pw, None, 2.0, 1, "lb", 0)
    monkeypatch.setenv("IGMK_DYNAMIC_BLOCKS", "1")
    with _engine(pop) as eng:
        dyn = eng.actdist(ii, jj, pw, None, 2.0, 1, "lb", 0)
    assert dyn.tobytes() == base.tobytes()


@pytest.mark.parametrize("nstruct", [100, 130, 700, 1000, 1024])
def test_list_form_matches_key_arrays_warp(nstruct, monkeypatch):
    """Small target probabilities (the sigma = 0.01 regime): the list-form kernel answers
    most pairs itself; results are byte-identical to the key-array kernels and the oracle."""
    from igm_b200 import synthetic
    pop = synthetic.make_population(2_000_000, nstruct, seed=300 + nstruct, genome_scale=0.05)
    rng = np.random.default_rng(nstruct)
    ii, jj = _sorted_pairs(rng, pop.n_hap, 40000)
    pw = np.exp(rng.uniform(np.log(0.004), np.log(0.08), len(ii))).astype(np.float32).astype(np.float64)
    pw[::97] = rng.uniform(0.2, 1.0, len(pw[::97]))          # a few heavy pairs in between
    pl = np.where(rng.random(len(ii)) < 0.5, 0.0,
                  orc.text_roundtrip(rng.uniform(0, 0.05, len(ii))).astype(np.float64))
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("IGMK_LIST", flag)
        with _engine(pop) as eng:
            for mode in ("lb", "gp"):
                for it_corr in (0, 1):
                    out[flag, mode, it_corr] = eng.actdist(ii, jj, pw, pl, 2.0, it_corr, mode, 0)
                    if flag == "1":
                        redo = eng.last_redo_count()
                        assert 0 <= redo <= len(ii), (mode, it_corr, redo)
    for mode in ("lb", "gp"):
        for it_corr in (0, 1):
            assert out["1", mode, it_corr].tobytes() == out["0", mode, it_corr].tobytes(), (mode, it_corr)
    sel = np.sort(rng.choice(len(ii), 400, replace=False))
    for mode in ("lb", "gp"):
        _, dets = orc.run_pairs(ii[sel], jj[sel], pw[sel], pl[sel], pop.coordinates, pop.radii, pop.chrom_hap(),
                                pop.copy_index, 1, 2.0, MODES[mode])
        _check_against_details(out["1", mode, 1][sel], dets)


@pytest.mark.parametrize("nstruct", [1500, 4100, 10000])
def test_list_form_matches_key_arrays_block(nstruct, monkeypatch):
    """Populations of more than 1024 structures: the slab pipeline (sample / fill / select
    kernels over slabs of 1024 structures), the one-CTA-per-pair list kernel (IGMK_SLAB=0)
    and the key-array kernels (IGMK_LIST=0) give byte-identical results."""
    from igm_b200 import synthetic
    pop = synthetic.make_population(2_000_000, nstruct, seed=17 + nstruct, genome_scale=0.01)
    rng = np.random.default_rng(nstruct)
    ii, jj = _sorted_pairs(rng, pop.n_hap, 3000)
    pw = np.exp(rng.uniform(np.log(0.004), np.log(0.08), len(ii))).astype(np.float32).astype(np.float64)
    pw[::53] = rng.uniform(0.2, 1.0, len(pw[::53]))
    pl = np.zeros(len(ii))
    out = {}
    for flag, slab in (("1", "1"), ("1", "0"), ("0", "1")):
        monkeypatch.setenv("IGMK_LIST", flag)
        monkeypatch.setenv("IGMK_SLAB", slab)
        with _engine(pop) as eng:
            for mode in ("lb", "gp"):
                out[flag, slab, mode] = eng.actdist(ii, jj, pw, pl, 2.0, 1, mode, 0)
                if flag == "1":
                    assert 0 <= eng.last_redo_count() <= len(ii)
    for mode in ("lb", "gp"):
        assert out["1", "1", mode].tobytes() == out["0", "1", mode].tobytes(), mode
        assert out["1", "0", mode].tobytes() == out["0", "1", mode].tobytes(), mode
    sel = np.sort(rng.choice(len(ii), 60, replace=False))
    _, dets = orc.run_pairs(ii[sel], jj[sel], pw[sel], pl[sel], pop.coordinates, pop.radii, pop.chrom_hap(),
                            pop.copy_index, 1, 2.0, MODES["lb"])
    _check_against_details(out["1", "1", "lb"][sel], dets)


@pytest.mark.parametrize("kind", ["near_first", "far_first", "grid"])
def test_list_form_unrepresentative_sample(kind, monkeypatch):
    """The list form takes its threshold from the first 128 structures.  Populations whose
    first structures are NOT representative (all compact / all extended) make it misjudge
    the threshold - too few values kept, or thread lists overflowing - and coordinates on a
    coarse grid give thousands of tied distances.  The answers must not change."""
    from igm_b200 import synthetic
    from igm_b200.population import Population
    nstruct = 900
    pop0 = synthetic.make_population(2_000_000, nstruct, seed=77, genome_scale=0.03)
    crd = pop0.coordinates.copy()
    if kind == "grid":
        crd = (np.round(crd / 400.0) * 400.0).astype(np.float32)
    else:
        scale = np.ones(nstruct, np.float32)
        scale[:128] = 0.05 if kind == "near_first" else 1.0
        scale[128:] = 1.0 if kind == "near_first" else 0.05
        crd = crd * scale[None, :, None]
    pop = Population(crd, pop0.radii, pop0.chrom, pop0.copy_index, pop0.copy)
    rng = np.random.default_rng(5)
    ii, jj = _sorted_pairs(rng, pop.n_hap, 6000)
    pw = np.exp(rng.uniform(np.log(0.004), np.log(0.3), len(ii))).astype(np.float32).astype(np.float64)
    pl = np.zeros(len(ii))
    with _engine(pop) as eng:
        res = {m: eng.actdist(ii, jj, pw, pl, 2.0, 1, m, 0) for m in ("lb", "gp")}
    monkeypatch.setenv("IGMK_LIST", "0")
    with _engine(pop) as eng:
        for m in ("lb", "gp"):
            assert eng.actdist(ii, jj, pw, pl, 2.0, 1, m, 0).tobytes() == res[m].tobytes(), m
    sel = np.sort(rng.choice(len(ii), 300, replace=False))
    for m in ("lb", "gp"):
        _, dets = orc.run_pairs(ii[sel], jj[sel], pw[sel], pl[sel], pop.coordinates, pop.radii, pop.chrom_hap(),
                                pop.copy_index, 1, 2.0, MODES[m])
        _check_against_details(res[m][sel], dets)


def test_degenerate_inputs():
    """Identical coordinates (all distances equal), empty and i == j inputs."""
    from igm_b200.population import CopyIndex, Population
    from igm_b200 import synthetic
    nstruct = 300
    bins = np.array([6, 5, 3])
    chrom_hap, chrom_bead, copy_bead, ci = synthetic.build_index(bins, n_diploid_chroms=2)
    nbead = len(chrom_bead)
    crd = np.zeros((nbead, nstruct, 3), np.float32)
    crd[:, :, 0] = np.arange(nbead, dtype=np.float32)[:, None] * 100.0   # same in every structure
    pop = Population(crd, np.full(nbead, 50.0, np.float32), chrom_bead, ci, copy_bead)
    ii = np.array([0, 0, 1, 2, 7, 3, 3], np.int32)
    jj = np.array([1, 7, 12, 13, 12, 3, 9], np.int32)    # (3,3): i == j
    pw = np.array([1.0, 0.3, 0.2, 0.01, 0.5, 0.5, 0.999], np.float64)
    pl = np.zeros(len(ii))
    with _engine(pop) as eng:
        for mode in ("lb", "gp"):
            _, dets = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                    pop.copy_index, 0, 2.0, MODES[mode])
            for algo in (0, 1):
                res = eng.actdist(ii, jj, pw, pl, 2.0, 0, mode, algo)
                _check_against_details(res, dets)
        # empty input
        res = eng.actdist(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), np.zeros(0))
        assert len(res) == 0
        row, col, dist, prob = eng.expand_records(np.zeros(0, np.int32), np.zeros(0, np.int32), res)
        assert len(row) == 0


def test_tiny_and_denormal_distances():
    """Keys in the bf16-denormal range and exact zeros."""
    from igm_b200.population import CopyIndex, Population
    from igm_b200 import synthetic
    nstruct = 200
    bins = np.array([4, 4])
    chrom_hap, chrom_bead, copy_bead, ci = synthetic.build_index(bins, n_diploid_chroms=2)
    nbead = len(chrom_bead)
    rng = np.random.default_rng(5)
    crd = (rng.standard_normal((nbead, nstruct, 3)) * 1e-20).astype(np.float32)
    crd[:, ::7, :] = 0.0
    pop = Population(crd, np.full(nbead, 1e-21, np.float32), chrom_bead, ci, copy_bead)
    ii, jj = np.triu_indices(8, 1)
    ii, jj = ii.astype(np.int32), jj.astype(np.int32)
    pw = rng.uniform(0.01, 1.0, len(ii))
    pl = np.zeros(len(ii))
    with _engine(pop) as eng:
        _, dets = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                pop.copy_index, 1, 2.0, orc.MODE_LB)
        res = eng.actdist(ii, jj, pw, pl, 2.0, 1, "lb", 0)
        _check_against_details(res, dets)


@pytest.mark.skipif(not H.have_demo(), reason="oracle/_ref/demo not present")
@pytest.mark.parametrize("sigma", ["1", "0.2", "0.1", "0.05", "0.02", "0.01"])
def test_full_demo_sha(sigma):
    """Config 1: the shipped demo population, every sigma of the demo config;
    the '%d.out.tmp' text must hash to what the reference's get_actdist gave."""
    from igm_b200.population import Population, ProbMatrix
    pop = Population.from_hss(H.DEMO_HSS)
    pm = ProbMatrix.from_hcs(H.DEMO_HCS)
    s = H.demo_full_summary()["sigmas"][sigma]
    ci, cj, cp = orc.select_candidates(pm.indptr, pm.indices, pm.data, pm.chrom,
                                       float(sigma), float(sigma), "float64")
    with _engine(pop) as eng:
        for mode in ("lb", "gp"):
            if mode not in s:
                continue
            res = eng.actdist(ci, cj, cp, None, 2.0, 0, mode)
            row, col, dist, prob = eng.expand_records(ci, cj, res)
            assert len(row) == s[mode]["records"]
            ad = np.sqrt(res["d2_sel_bits"].view(np.float32).astype(np.float64))
            ad_r = np.repeat(ad, res["nrec"])
            p_r = np.repeat(res["p"], res["nrec"])
            text = "\n".join(["%6d %6d %10.4f %.4f" % x for x in zip(row.tolist(), col.tolist(), ad_r.tolist(), p_r.tolist())])
            assert hashlib.sha256(text.encode()).hexdigest() == s[mode]["sha256"]
            # the device-side 4-decimal rounding equals the text round trip
            if sigma != "0.01":
                assert np.array_equal(dist.view(np.uint32), orc.text_roundtrip(ad_r).view(np.uint32))
                assert np.array_equal(prob.view(np.uint32), orc.text_roundtrip(p_r).view(np.uint32))


def test_mixed_ploidy_inside_a_chromosome_is_refused_in_lb_mode():
    """the reference's LB intra branch reads uninitialised memory when the two loci
    of one chromosome have different copy counts; the engine refuses such pairs (GP mode and
    inter-chromosomal pairs are well defined and still match the oracle)."""
    from igm_b200.population import CopyIndex, Population
    rng = np.random.default_rng(12)
    # loci 0..3 on chromosome 0 (locus 1 has a single copy), loci 4..5 on chromosome 1
    ptr = np.array([0, 2, 3, 5, 7, 9, 10], np.int32)
    beads = np.array([0, 6, 1, 2, 7, 3, 8, 4, 9, 5], np.int32)
    chrom_hap = np.array([0, 0, 0, 0, 1, 1], np.int32)
    nbead, nstruct = 10, 64
    crd = (rng.standard_normal((nbead, nstruct, 3)) * 400).astype(np.float32)
    chrom_bead = np.zeros(nbead, np.int32)
    for h in range(6):
        chrom_bead[beads[ptr[h]:ptr[h + 1]]] = chrom_hap[h]
    pop = Population(crd, np.full(nbead, 60.0, np.float32), chrom_bead, CopyIndex(ptr, beads))
    with _engine(pop) as eng:
        with pytest.raises(ValueError):
            eng.actdist(np.array([0], np.int32), np.array([1], np.int32), np.array([0.5]), None, 2.0, 0, "lb")
        ii = np.array([0, 0, 1, 1, 2], np.int32)
        jj = np.array([2, 4, 4, 5, 5], np.int32)          # equal-count intra pair and inter pairs
        pw = np.array([0.5, 0.2, 0.9, 0.05, 1.0])
        for mode in ("lb", "gp"):
            _, dets = orc.run_pairs(ii, jj, pw, np.zeros(5), pop.coordinates, pop.radii, pop.chrom_hap(),
                                    pop.copy_index, 0, 2.0, MODES[mode])
            _check_against_details(eng.actdist(ii, jj, pw, None, 2.0, 0, mode), dets)
        _, dets = orc.run_pairs(np.array([0], np.int32), np.array([1], np.int32), np.array([0.5]), np.zeros(1),
                                pop.coordinates, pop.radii, pop.chrom_hap(), pop.copy_index, 0, 2.0, orc.MODE_GP)
        _check_against_details(eng.actdist(np.array([0], np.int32), np.array([1], np.int32), np.array([0.5]),
                                           None, 2.0, 0, "gp"), dets)


def test_round4_random():
    """round4_to_f32 on device == float32(float('%.4f' % x)) incl. exact ties."""
    from igm_b200.population import Population
    from igm_b200 import synthetic
    # drive it through prob: p = pwish (it_corr = 0), any positive float64
    pop = synthetic.make_population(2_000_000, 16, seed=3, genome_scale=0.004)
    rng = np.random.default_rng(11)
    n = 4000
    pw = np.concatenate([rng.uniform(0, 1, n), np.arange(1, 400, 2) / 32.0 / 16.0,
                         np.array([0.00005, 0.00015, 0.03125, 0.5 + 1 / 32.0, 1e-9, 0.99995, 1.0])])
    ii = np.zeros(len(pw), np.int32)
    jj = np.ones(len(pw), np.int32)
    with _engine(pop) as eng:
        res = eng.actdist(ii, jj, pw, None, 2.0, 0, "lb")
    assert np.array_equal(res["prob"].view(np.uint32), orc.text_roundtrip(pw).view(np.uint32))
    ad = np.sqrt(res["d2_sel_bits"].view(np.float32).astype(np.float64))
    assert np.array_equal(res["dist"].view(np.uint32), orc.text_roundtrip(ad).view(np.uint32))


def test_device_entry_point_and_errors():
    import torch
    from igm_b200 import synthetic, _lib
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, 100, seed=21, genome_scale=0.01)
    rng = np.random.default_rng(2)
    ii = rng.integers(0, pop.n_hap - 1, 500).astype(np.int32)
    jj = (ii + 1 + rng.integers(0, 3, 500)).clip(max=pop.n_hap - 1).astype(np.int32)
    pw = rng.uniform(0.01, 1, 500)
    with ActdistEngine(pop, 0) as eng:
        host = eng.actdist(ii, jj, pw)
        dev = torch.device("cuda:0")
        d_i, d_j = torch.from_numpy(ii).to(dev), torch.from_numpy(jj).to(dev)
        d_pw, d_pl = torch.from_numpy(pw).to(dev), torch.zeros(500, dtype=torch.float64, device=dev)
        d_out = torch.zeros(500 * 32, dtype=torch.uint8, device=dev)
        eng.actdist_device(d_i, d_j, d_pw, d_pl, d_out, 500)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(_lib.PAIR_RESULT_DTYPE)
        assert got.tobytes() == host.tobytes()
        # coordinates staged from a CUDA tensor give the same answer
        eng.upload_coordinates(torch.from_numpy(pop.coordinates).to(dev))
        assert eng.actdist(ii, jj, pw).tobytes() == host.tobytes()
        with pytest.raises(ValueError):
            eng.actdist(np.array([0], np.int32), np.array([pop.n_hap], np.int32), np.array([0.5]))
    with pytest.raises(_lib.IgmkError):
        ActdistEngine(nbead=10, nstruct=10, device=99)


@pytest.mark.parametrize("nstruct", [300, 2600])
def test_peer_store_entry_point_single_gpu(nstruct):
    """igmk_actdist_device_peers with the 'peers' being two buffers on this one GPU
    (warp and CTA groups): every copy equals the ordinary path after the finish pass."""
    import torch
    from igm_b200 import synthetic, _lib
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, nstruct, seed=77, genome_scale=0.01)
    rng = np.random.default_rng(4)
    n = 3000
    ii = rng.integers(0, pop.n_hap - 1, n).astype(np.int32)
    jj = (ii + 1 + rng.integers(0, 9, n)).clip(max=pop.n_hap - 1).astype(np.int32)
    nc, ch = pop.copy_index.ncopies(), pop.chrom_hap()
    bad = ((ch[ii] == ch[jj]) & (nc[ii] != nc[jj]))
    jj[bad] = ii[bad]                                         # i == j: empty result slot
    pw = rng.uniform(0.005, 1, n)
    dev = torch.device("cuda:0")
    with ActdistEngine(pop, 0) as eng:
        ref = eng.actdist(ii, jj, pw, None, 2.0, 1, "LB")
        d_i, d_j = torch.from_numpy(ii).to(dev), torch.from_numpy(jj).to(dev)
        d_pw, d_pl = torch.from_numpy(pw).to(dev), torch.zeros(n, dtype=torch.float64, device=dev)
        bufs = torch.zeros((2, n, 32), dtype=torch.uint8, device=dev)
        peers = torch.tensor([bufs[0].data_ptr(), bufs[1].data_ptr()], dtype=torch.int64, device=dev)
        eng.actdist_device_peers(d_i, d_j, d_pw, d_pl, peers, 2, n, 2.0, 1, "LB")
        eng.finish_results(bufs, 2 * n)
        torch.cuda.synchronize()
        got = bufs.cpu().numpy().reshape(2, -1).view(_lib.PAIR_RESULT_DTYPE)
    assert got[0].tobytes() == ref.tobytes() and got[1].tobytes() == ref.tobytes()


@pytest.mark.parametrize("n_peers", [1, 3, 8, 11])
@pytest.mark.parametrize("n", [5, 1000, 60001])
def test_peer_store_many_buffers(n_peers, n):
    """Multi-GPU emit (finished records stored into every gather buffer from inside K1) with 1 to
    11 buffers, all on this GPU: every buffer equals the ordinary path - short and long lists,
    pairs the list form hands to the key-array kernel, invalid pairs in between."""
    import torch
    from igm_b200 import synthetic, _lib
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, 700, seed=5, genome_scale=0.03)
    rng = np.random.default_rng(n + n_peers)
    ii = rng.integers(0, pop.n_hap - 1, n).astype(np.int32)
    jj = (ii + 1 + rng.integers(0, 40, n)).clip(max=pop.n_hap - 1).astype(np.int32)
    nc, ch = pop.copy_index.ncopies(), pop.chrom_hap()
    bad = ((ch[ii] == ch[jj]) & (nc[ii] != nc[jj]))
    jj[bad] = ii[bad]                                         # i == j: empty result slot
    pw = np.exp(rng.uniform(np.log(0.004), np.log(0.2), n))
    pw[::17] = 0.9                                            # heavy pairs: handed to the key-array kernel
    dev = torch.device("cuda:0")
    with ActdistEngine(pop, 0) as eng:
        ref = eng.actdist(ii, jj, pw, None, 2.0, 1, "LB")
        d_i, d_j = torch.from_numpy(ii).to(dev), torch.from_numpy(jj).to(dev)
        d_pw, d_pl = torch.from_numpy(pw).to(dev), torch.zeros(n, dtype=torch.float64, device=dev)
        bufs = torch.full((n_peers, n, 32), 0xEE, dtype=torch.uint8, device=dev)
        peers = torch.tensor([bufs[k].data_ptr() for k in range(n_peers)], dtype=torch.int64, device=dev)
        eng.actdist_device_peers(d_i, d_j, d_pw, d_pl, peers, n_peers, n, 2.0, 1, "LB")
        torch.cuda.synchronize()
        got = bufs.cpu().numpy().reshape(n_peers, -1).view(_lib.PAIR_RESULT_DTYPE)
    for k in range(n_peers):
        assert got[k].tobytes() == ref.tobytes(), k


def test_swapped_and_duplicate_pairs():
    """i > j pairs, duplicates and unsorted order behave like independent get_actdist calls."""
    from igm_b200 import synthetic
    pop = synthetic.make_population(2_000_000, 150, seed=5, genome_scale=0.02)
    rng = np.random.default_rng(9)
    nh = pop.n_hap
    ii = rng.integers(0, nh, 400)
    jj = rng.integers(0, nh, 400)
    nc, ch = pop.copy_index.ncopies(), pop.chrom_hap()
    ok = (ii != jj) & ~((ch[ii] == ch[jj]) & (nc[ii] != nc[jj]))
    ii, jj = ii[ok].astype(np.int32), jj[ok].astype(np.int32)          # both orders occur
    ii = np.concatenate([ii, ii[:50], jj[:50]])                        # duplicates and swapped duplicates
    jj = np.concatenate([jj, jj[:50], ii[:50]])
    assert (ii > jj).any() and (ii < jj).any()
    pw = rng.uniform(0.001, 1.0, len(ii)).astype(np.float32).astype(np.float64)
    pl = np.zeros(len(ii))
    with _engine(pop) as eng:
        for mode in ("lb", "gp"):
            recs, dets = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                       pop.copy_index, 1, 2.0, MODES[mode])
            res = eng.actdist(ii, jj, pw, pl, 2.0, 1, mode, 0)
            _check_against_details(res, dets)
            row, col, dist, prob = eng.expand_records(ii, jj, res)
            erow, ecol, edist, eprob = orc.records_to_arrays(recs)
            assert np.array_equal(row, erow) and np.array_equal(col, ecol)
            assert np.array_equal(dist, edist) and np.array_equal(prob, eprob)


@pytest.mark.parametrize("nstruct", [100, 1000, 2500])
def test_sel_flat_idx(nstruct):
    """Selected structure index: d_sq[0:npc].ravel()[sel_flat_idx] is the
    selected value, bit for bit, and the index is the lowest one holding that value."""
    from igm_b200 import synthetic
    pop = synthetic.make_population(2_000_000, nstruct, seed=900 + nstruct, genome_scale=0.02)
    crd = pop.coordinates.copy()
    crd[:, : nstruct // 3] = np.round(crd[:, : nstruct // 3] / 300.0) * 300.0      # ties among structures
    from igm_b200.population import Population
    pop = Population(crd, pop.radii, pop.chrom, pop.copy_index, pop.copy)
    rng = np.random.default_rng(nstruct)
    ii, jj = _sorted_pairs(rng, pop.n_hap, 260)
    pw = rng.uniform(0.002, 1.0, len(ii)).astype(np.float32).astype(np.float64)
    pw[::7] = 0.0                                            # no record -> -1
    ch, ci = pop.chrom_hap(), pop.copy_index
    with _engine(pop) as eng:
        for mode in ("lb", "gp"):
            res = eng.actdist(ii, jj, pw, None, 2.0, 0, mode, 0)
            idx = eng.sel_flat_idx(ii, jj, res, mode)
            for t in range(len(ii)):
                if res["o"][t] < 0:
                    assert idx[t] == -1
                    continue
                a, b = ci[int(ii[t])], ci[int(jj[t])]
                if mode == "lb" and ch[ii[t]] == ch[jj[t]]:
                    combos, npc = list(zip(a, b)), min(len(a), len(b))
                else:
                    combos = [(k, m) for k in a for m in b]
                    npc = len(a) * len(b) if mode == "lb" else min(len(a), len(b))
                d_sq = np.stack([np.sum(np.square(crd[k] - crd[m]), axis=1) for k, m in combos]).astype(np.float64)
                d_sq.sort(axis=0)                                                  # :439
                flat = d_sq[0:npc].ravel().astype(np.float32).view(np.uint32)
                assert flat[idx[t]] == res["d2_sel_bits"][t], (mode, t)
                assert idx[t] == np.flatnonzero(flat == res["d2_sel_bits"][t])[0], (mode, t)
