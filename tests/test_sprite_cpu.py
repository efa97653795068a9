"""K4 oracles: the compiled reference (oracle/_ref/libsprite_ref.so, built from the
unmodified /root/reference source) against the known answers of
igm/cython_compiled/tests.py, and the NumPy port against the compiled reference."""
import numpy as np
import pytest

from oracle import sprite_oracle as so

KATS = [
    (np.array([[[1, 0, 0]], [[-1, 0, 0]], [[0, -1, 0]], [[0, 1, 0]]], np.float32), [1, 1, 1, 1],
     ([1.0], 0, [[0, 0, 0, 0]])),
    (np.array([[[1, 0, 0]], [[0.5, 0, 0]], [[0, -0.5, 0]], [[0, 0.5, 0]]], np.float32), [2, 2],
     ([0.125], 0, [[1, 0]])),                       # executed value (the comment in tests.py is stale)
    (np.array([[[1, 0, 0], [0.1, 0, 0]], [[0.5, 0, 0], [1, 0, 0]], [[0, -0.5, 0], [-1, 0, 0]],
               [[0, 0.5, 0], [-0.1, 0, 0]]], np.float32), [2, 2],
     ([0.125, 0.01], 1, [[1, 0], [0, 1]])),
]


@pytest.mark.parametrize("case", range(3))
def test_known_answers(case):
    crd, cn, (rg, best, ci) = KATS[case]
    fns = [so.get_rgs2_port] + ([so.ref_get_rgs2] if so.ref_available() else [])
    for fn in fns:
        r, b, c = fn(crd, np.array(cn, np.int32))
        assert np.allclose(r, rg, rtol=1e-6) and b == best and c.tolist() == ci


@pytest.mark.skipif(not so.ref_available(), reason="oracle/_ref/libsprite_ref.so not built")
def test_port_equals_compiled_reference():
    rng = np.random.default_rng(0)
    for copies in ([2, 2, 2], [1, 2, 1, 2], [2], [2, 2, 2, 2, 2], [3, 1, 2]):
        b = int(np.sum(copies))
        crd = (rng.standard_normal((b, 57, 3)) * 1500).astype(np.float32)
        crd[:, 5] = crd[:, 4]                       # ties between structures
        r1, b1, c1 = so.ref_get_rgs2(crd, np.array(copies, np.int32))
        r2, b2, c2 = so.get_rgs2_port(crd, copies)
        assert np.array_equal(r1.view(np.uint32), r2.view(np.uint32))
        assert b1 == b2 and np.array_equal(c1, c2)


def test_compute_gyration_radius_port_equals_reference_golden():
    """The cluster-level driver against what the reference's own compiled
    compute_gyration_radius returned (tests/golden/make_golden_sprite.py), with the global
    NumPy random stream seeded as it was there; both get_rgs2 oracles."""
    from tests import helpers as H
    fns = [so.get_rgs2_port] + ([so.ref_get_rgs2] if so.ref_available() else [])
    n = 0
    for name, crd, chrom, ci, clusters, rg2s, best, sel, seed0, _ in H.sprite_golden_cases():
        for k, cl in enumerate(clusters):
            for fn in fns:
                np.random.seed(seed0 + k)
                r, b, s = so.compute_gyration_radius_port(crd, cl, chrom, ci, get_rgs2=fn)
                assert np.array_equal(np.asarray(r, np.float32).view(np.uint32), rg2s[k].view(np.uint32)), (name, k)
                assert int(b) == int(best[k])
                assert np.array_equal(np.asarray(s), sel[k])
            n += 1
    assert n == 54
