"""TEST INFRASTRUCTURE ONLY - CPU restatement of the Hi-C restraint selection of
the M-step: intraHiC._apply / interHiC._apply (igm/restraints/intra_hic.py:39-58,
inter_hic.py:39-58) with Particle.__sub__ (igm/model/particle.py:35-36).

Pinned: tests/golden/make_golden_restraint.py runs the reference's own classes
(imported through oracle/ref_loader.py; ActivationDistanceDB replaced by a list,
the model by a recorder of addForce calls) and stores the selected (i, j) lists in
tests/golden/restraint_small.npz.

np.linalg.norm is called exactly as the reference calls it, so the restatement
inherits the host BLAS's float32 dot arithmetic (float32 products accumulated in
float64, rounded once: OpenBLAS 0.3.30 sdot tail loop) - the arithmetic the CUDA
kernel reproduces explicitly (igm_b200/csrc/igmk_restraint.cuh).
"""
import numpy as np


def select_for_structure(pos, chrom, row, col, dist, kind="intra"):
    """pos: (nbead, 3) float32 coordinates of ONE structure.  Returns the indices of
    the records that get a bond, in record order."""
    out = []
    for k, (i, j, d) in enumerate(zip(row, col, dist)):
        same = chrom[i] == chrom[j]
        if kind == "intra" and not same:
            continue
        if kind == "inter" and same:
            continue
        if np.linalg.norm(pos[i] - pos[j]) <= d:
            out.append(k)
    return np.array(out, dtype=np.int64)


def select_bitmap(coords, chrom, row, col, dist, kind="intra"):
    """coords: (nbead, nstruct, 3).  bool matrix (n_rec, nstruct)."""
    n_rec, nstruct = len(row), coords.shape[1]
    out = np.zeros((n_rec, nstruct), dtype=bool)
    for s in range(nstruct):
        out[select_for_structure(coords[:, s, :], chrom, row, col, dist, kind), s] = True
    return out


def dot3_model(d):
    """The explicit arithmetic behind np.linalg.norm(d)**2 for float32 3-vectors in this
    image: float32 products, float64 accumulation, one rounding to float32."""
    d = np.asarray(d, dtype=np.float32)
    sq = (d * d).astype(np.float32)
    return (sq[..., 0].astype(np.float64) + sq[..., 1].astype(np.float64) + sq[..., 2].astype(np.float64)).astype(np.float32)
