"""Population contact-frequency maps on B200 (K2), host side.

Mirrors the two call shapes of the reference:

* ``alabtools.analysis.get_simulated_hic(hss, contact_range)`` as called by
  ``report_hic`` (igm/report/hic.py:49-52);
* ``HssFile.buildContactMap(contactRange=...)`` followed by
  ``Contactmatrix.sumCopies()`` (igm/steps/HicEvaluationStep.py:107-112).

``alabtools`` is not part of the reference tree, so the arithmetic follows the
reference's own in-tree definitions (A-step contact test, float32, inclusive;
``strict=True`` gives the ``<`` of HicEvaluationStep.py:89-92) and the copy
projection is the plain sum over copy combinations - PARITY UNPINNED against
alabtools.  All arithmetic runs in ``contact_tile_kernel``
(igm_b200/csrc/igmk_contact.cuh); there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .engine import ActdistEngine
from .population import Population, ProbMatrix


def row_blocks(n: int, block: int, rank: int = 0, world: int = 1):
    """Row blocks [r0, r1) of an n x n upper-triangular sweep owned by `rank`.
    Block b goes to rank b % world if (b // world) is even, else to
    world - 1 - b % world (boustrophedon): block rows of an upper triangle shrink
    linearly, so pairing long with short rows balances the ranks without any
    exchange (K2 tiles are independent, no collective)."""
    out = []
    nb = (n + block - 1) // block
    for b in range(nb):
        owner = b % world if (b // world) % 2 == 0 else world - 1 - b % world
        if owner == rank:
            out.append((b * block, min(n, (b + 1) * block)))
    return out


def haploid_contact_counts(eng: ActdistEngine, contact_range: float = 2.0, strict: bool = False,
                           block: int = 2048, rank: int = 0, world: int = 1,
                           out: Optional[np.ndarray] = None) -> np.ndarray:
    """Dense (n_hap, n_hap) uint32 matrix of copy-summed contact counts.  Only the
    block-upper-triangle is computed on the GPU (rows owned by `rank`); the lower
    triangle is mirrored on the host.  With world > 1 the rows of other ranks stay
    zero (combine with a sum / gather)."""
    n = eng.n_hap
    if out is None:
        out = np.zeros((n, n), dtype=np.uint32)
    for r0, r1 in row_blocks(n, block, rank, world):
        tile = eng.contact_counts_haploid(r0, r1 - r0, r0, n - r0, contact_range, strict)
        out[r0:r1, r0:] = tile
        out[r0:, r0:r1] = tile.T          # mirror (diagonal block written twice, identical)
    return out


def counts_to_probmatrix(counts: np.ndarray, nstruct: int, chrom_hap: np.ndarray,
                         clip: bool = False) -> ProbMatrix:
    """Strict-upper-triangle CSR of count / nstruct (float32), diagonal kept apart -
    the layout of a .hcs file."""
    n = counts.shape[0]
    iu, ju = np.nonzero(np.triu(counts, 1))
    data = (counts[iu, ju].astype(np.float64) / float(nstruct)).astype(np.float32)
    diag = (np.diagonal(counts).astype(np.float64) / float(nstruct)).astype(np.float32)
    if clip:
        data = data.clip(0, 1)
        diag = diag.clip(0, 1)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(iu, minlength=n), out=indptr[1:])
    return ProbMatrix(indptr, ju.astype(np.int32), data, chrom_hap, diag)


def get_simulated_hic(hss, contact_range: float = 2.0, strict: bool = False, device: int = 0,
                      block: int = 2048) -> ProbMatrix:
    """Haploid contact-probability matrix of a population (igm/report/hic.py:51).
    `hss`: path of a .hss file or a Population."""
    pop = hss if isinstance(hss, Population) else Population.from_hss(hss)
    with ActdistEngine(pop, device) as eng:
        counts = haploid_contact_counts(eng, contact_range, strict, block)
    return counts_to_probmatrix(counts, pop.nstruct, pop.chrom_hap())


def evaluation_stats(inp: ProbMatrix, out: ProbMatrix, sigma: float) -> Tuple[float, float, float]:
    """(score, mean difference, mean relative difference) of HicEvaluationStep.reduce
    (:156-176): over stored input pairs with pwish >= sigma and i != j that also
    have a stored output value."""
    n = inp.n
    ki = inp.rows().astype(np.int64) * n + inp.indices.astype(np.int64)
    keep = (inp.data >= np.float32(sigma)) & (inp.rows() != inp.indices)
    ki, pi = ki[keep], inp.data[keep].astype(np.float64)
    ko = out.rows().astype(np.int64) * n + out.indices.astype(np.int64)
    po = out.data.astype(np.float64)
    pos = np.searchsorted(ki, ko)
    pos[pos >= len(ki)] = 0
    hit = (len(ki) > 0) & (ki[pos] == ko) if len(ki) else np.zeros(len(ko), bool)
    diffs = po[hit] - pi[pos[hit]]
    rel = diffs / pi[pos[hit]]
    if len(diffs) == 0:
        return float("nan"), float("nan"), float("nan")
    return float(np.abs(rel).mean()), float(np.average(diffs)), float(np.average(rel))
