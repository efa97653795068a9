"""TEST INFRASTRUCTURE ONLY - CPU restatement (NumPy) of the population
contact-frequency counts.

PARITY UNPINNED: the arithmetic of ``HssFile.buildContactMap`` /
``get_simulated_hic`` lives in the third-party package ``alabtools``
(setup.py:19, ``alabtools>=0.0.1``, not vendored, no lockfile), so there is no
reference implementation or golden vector to check against.  This restatement
follows the reference's own in-tree definitions: the A-step's contact test
``d_sq <= np.square(contactRange * (ri + rj))`` in float32
(igm/steps/ActivationDistanceStep.py:396,418,442) and, with ``strict=True``,
the ``<`` of the commented block igm/steps/HicEvaluationStep.py:73-93.
"""
import numpy as np


def contact_counts(coords, radii, rows, cols, contact_range=2.0, strict=False):
    """counts[a, b] = #{s : d2_s(a, b) <= (cr (r_a + r_b))^2} for a in rows, b in cols."""
    rows = np.asarray(rows)
    cols = np.asarray(cols)
    out = np.zeros((len(rows), len(cols)), dtype=np.uint32)
    cr = np.float32(contact_range)
    for ia, a in enumerate(rows):
        x = coords[a]
        for ib, b in enumerate(cols):
            y = coords[b]
            d_sq = np.sum(np.square(x - y), axis=1)
            rcutsq = np.square(cr * (radii[a] + radii[b]))
            out[ia, ib] = np.count_nonzero(d_sq < rcutsq if strict else d_sq <= rcutsq)
    return out


def contact_counts_fast(coords, radii, rows, cols, contact_range=2.0, strict=False):
    """Vectorised over columns; same float32 operation order per element."""
    rows = np.asarray(rows)
    cols = np.asarray(cols)
    cr = np.float32(contact_range)
    out = np.zeros((len(rows), len(cols)), dtype=np.uint32)
    Y = coords[cols]                                   # (nc, N, 3)
    for ia, a in enumerate(rows):
        diff = coords[a][None, :, :] - Y               # float32
        sq = np.square(diff)
        d_sq = (sq[..., 0] + sq[..., 1]) + sq[..., 2]  # sequential float32 sum
        rcutsq = np.square(cr * (radii[a] + radii[cols]))[:, None]
        hit = d_sq < rcutsq if strict else d_sq <= rcutsq
        out[ia] = hit.sum(axis=1)
    return out


def sum_copies(counts_full, copy_ptr, copy_beads):
    """Haploid projection: f_IJ = sum over copies a of I, b of J of f_ab
    (Contactmatrix.sumCopies as used at HicEvaluationStep.py:111; semantics
    restated, see PARITY UNPINNED above)."""
    n_hap = len(copy_ptr) - 1
    nbead = counts_full.shape[0]
    P = np.zeros((n_hap, nbead), dtype=np.float64)
    for i in range(n_hap):
        P[i, copy_beads[copy_ptr[i]:copy_ptr[i + 1]]] = 1.0
    return P @ counts_full.astype(np.float64) @ P.T
