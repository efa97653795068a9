"""The NumPy oracle against the golden vectors produced by the reference's own
get_actdist (tests/golden/make_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import actdist_oracle as orc
from tests import helpers as H

MODES = {"lb": orc.MODE_LB, "gp": orc.MODE_GP}


@pytest.mark.parametrize("case", [c[0] for c in H.all_small_cases()])
@pytest.mark.parametrize("mode", ["lb", "gp"])
@pytest.mark.parametrize("it_corr", [0, 1])
def test_oracle_matches_reference_small(case, mode, it_corr):
    name, npz, prefix = [c for c in H.all_small_cases() if c[0] == case][0]
    pop, ii, jj, pw, pl = H.load_case(npz, prefix)
    g = H.golden_out(npz, prefix, mode, it_corr)
    recs, dets = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                               pop.copy_index, it_corr, 2.0, MODES[mode])
    assert [d.nrec for d in dets] == list(g["nrec"])
    assert np.array_equal(np.array([r[0] for r in recs], np.int32), g["row"])
    assert np.array_equal(np.array([r[1] for r in recs], np.int32), g["col"])
    # bit-exact float64 activation distances and probabilities
    assert np.array_equal(np.array([float(r[2]) for r in recs]).view(np.uint64), g["ad"].view(np.uint64))
    assert np.array_equal(np.array([float(r[3]) for r in recs]).view(np.uint64), g["p"].view(np.uint64))
    assert hashlib.sha256(orc.task_text(recs).encode()).hexdigest() == g["sha"]


def test_oracle_contact_range_variant():
    s = np.load(os.path.join(H.GOLDEN, "synth_small.npz"))
    pop, ii, jj, pw, pl = H.load_case(s, "n100")
    for mode in ("lb", "gp"):
        for it in (0, 1):
            g = H.golden_out(s, "n100cr35", mode, it)
            recs, _ = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                    pop.copy_index, it, 3.5, MODES[mode])
            assert hashlib.sha256(orc.task_text(recs).encode()).hexdigest() == g["sha"]


def test_utils_actdist_equals_gp_itcorr1():
    # igm/utils/actdist.py has no it_corr argument and always corrects (:83-84)
    d = np.load(os.path.join(H.GOLDEN, "demo_subset.npz"))
    g = H.golden_out(d, "demo", "gp", 1)
    assert np.array_equal(d["demo_utils_ad"].view(np.uint64), g["ad"].view(np.uint64))
    assert np.array_equal(d["demo_utils_p"].view(np.uint64), g["p"].view(np.uint64))


def test_known_answers_from_survey():
    """first lines of the demo sigma=0.1 output and the full-run
    hashes (independent probe of the same reference function)."""
    s = H.demo_full_summary()["sigmas"]
    expect = {"1": ("e3c9a279f0163252", 1702, 3294), "0.2": ("a1af8846500f4d0e", 2824, 5444),
              "0.1": ("66afb2fcf0bf6f03", 6307, 12159), "0.05": ("c3b8ea46ecc7bbf7", 14994, 29373),
              "0.02": ("15396a7943992988", 44944, 104090), "0.01": ("83addf4d2a91e92b", 614139, 2287235)}
    for k, (sha, npairs, nrec) in expect.items():
        assert s[k]["lb"]["sha256"][:16] == sha
        assert s[k]["pairs_f64_compare"] == npairs
        assert s[k]["lb"]["records"] == nrec
    assert s["0.05"]["gp"]["sha256"][:16] == "e26e687dcc57ccda"
    assert s["0.02"]["gp"]["sha256"][:16] == "c6768bf7fac63060"
    # float32 comparison keeps one more pair at sigma = 0.01 (SURVEY 'filter edge')
    assert s["0.01"]["pairs_f32_compare"] == 614140


@pytest.mark.skipif(not H.have_demo(), reason="oracle/_ref/demo not built")
@pytest.mark.parametrize("sigma", ["1", "0.2", "0.1", "0.05"])
def test_oracle_full_demo_sigma(sigma):
    from igm_b200.population import Population, ProbMatrix
    pop = Population.from_hss(H.DEMO_HSS)
    pm = ProbMatrix.from_hcs(H.DEMO_HCS)
    s = H.demo_full_summary()["sigmas"][sigma]
    ci, cj, cp = orc.select_candidates(pm.indptr, pm.indices, pm.data, pm.chrom,
                                       float(sigma), float(sigma), "float64")
    assert len(ci) == s["pairs_f64_compare"]
    for mode in ("lb", "gp"):
        recs, _ = orc.run_pairs(ci, cj, cp, np.zeros(len(ci)), pop.coordinates, pop.radii,
                                pop.chrom_hap(), pop.copy_index, 0, 2.0, MODES[mode])
        assert hashlib.sha256(orc.task_text(recs).encode()).hexdigest() == s[mode]["sha256"]


def test_text_roundtrip_quirks():
    # SURVEY q4
    assert "%.4f" % 0.00005 == "0.0001"
    assert orc.text_roundtrip([0.00004])[0] == np.float32(0.0)
    assert orc.text_roundtrip([0.03125])[0] == np.float32(0.0312)   # exact tie -> half-even
    assert orc.text_roundtrip([1123.12044])[0] == np.float32(1123.1204)


@pytest.mark.parametrize("case", [c[0] for c in H.all_small_cases()])
def test_c_oracle_matches_numpy_oracle_and_golden(case):
    """oracle/actdist_oracle.c (plain C restatement) against the NumPy oracle,
    itself pinned to the reference's golden vectors above."""
    from oracle import c_oracle
    if not c_oracle.available():
        pytest.skip("run `make -C oracle`")
    name, npz, prefix = [c for c in H.all_small_cases() if c[0] == case][0]
    pop, ii, jj, pw, pl = H.load_case(npz, prefix)
    for mode in ("lb", "gp"):
        for it_corr in (0, 1):
            _, dets = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                    pop.copy_index, it_corr, 2.0, MODES[mode])
            exp = orc.details_to_arrays(dets)
            got = c_oracle.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                     pop.copy_index.ptr, pop.copy_index.beads, it_corr, 2.0, MODES[mode])
            for k in ("contact_count", "o", "nrec"):
                assert np.array_equal(got[k], exp[k]), (mode, it_corr, k)
            assert np.array_equal(got["p"].view(np.uint64), exp["p"].view(np.uint64))
            sel = exp["o"] >= 0
            assert np.array_equal(got["d2_sel_bits"][sel], exp["d2_sel_bits"][sel])
            g = H.golden_out(npz, prefix, mode, it_corr)
            assert np.array_equal(got["nrec"], g["nrec"])


def test_c_contact_oracle_matches_numpy():
    from oracle import c_oracle, contact_oracle as co
    if not c_oracle.available():
        pytest.skip("run `make -C oracle`")
    d = np.load(os.path.join(H.GOLDEN, "demo_subset.npz"))
    pop, *_ = H.load_case(d, "demo")
    rows, cols = np.arange(0, 40), np.arange(20, 70)
    for strict in (False, True):
        a = co.contact_counts_fast(pop.coordinates, pop.radii, rows, cols, 2.0, strict)
        b = c_oracle.contact_counts(pop.coordinates, pop.radii, rows, cols, 2.0, strict)
        assert np.array_equal(a, b)
