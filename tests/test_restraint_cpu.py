"""K3 oracle pinned to the reference's own intraHiC / interHiC._apply (golden vectors
from tests/golden/make_golden_restraint.py) and the explicit float32 dot model."""
import os

import numpy as np
import pytest

from oracle import restraint_oracle as ro

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "restraint_small.npz")


@pytest.mark.parametrize("kind", ["intra", "inter"])
def test_oracle_matches_reference_golden(kind):
    g = np.load(GOLDEN)
    coords = g["coords"]
    exp = np.unpackbits(g["sel_" + kind], axis=1)[:, :coords.shape[1]].astype(bool)
    got = ro.select_bitmap(coords, g["chrom"], g["row"], g["col"], g["dist"], kind)
    assert np.array_equal(got, exp)
    assert exp.any() and not exp.all()


def test_dot_model_equals_host_norm():
    """np.linalg.norm(float32 3-vector) == sqrt(float32(float64 sum of float32 squares))
    in this image - the arithmetic the CUDA kernel implements."""
    rng = np.random.default_rng(0)
    d = ((rng.standard_normal((20000, 3)) - rng.standard_normal((20000, 3))) * 2000).astype(np.float32)
    ref = np.array([np.linalg.norm(v) for v in d])
    assert ref.dtype == np.float32
    assert np.array_equal(np.sqrt(ro.dot3_model(d)), ref)
    tiny = (rng.standard_normal((2000, 3)) * 1e-18).astype(np.float32)
    assert np.array_equal(np.sqrt(ro.dot3_model(tiny)), np.array([np.linalg.norm(v) for v in tiny]))
