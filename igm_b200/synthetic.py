"""Seeded synthetic populations and Hi-C probability matrices.

Shapes follow the reference's own preprocessing (igm/_preprocess.py:15-110,
153-162): male diploid hg38 genome, 22 autosomes x 2 copies + X + Y, bins of
``resolution`` bp, ``copy_index[i] = [i, i + n_hap]`` for autosomal bins and
``[i]`` for X/Y (as in demo/demo_sample_outputs/igm-model.hss.T), uniform bead
radius, coordinates = per-structure confined random walk per chromosome copy.

There is no network here, so benchmarks and most tests run on these.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .population import CopyIndex, Population, ProbMatrix

HG38_LENGTHS = np.array([
    248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973,
    145138636, 138394717, 133797422, 135086622, 133275309, 114364328, 107043718,
    101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983,
    50818468, 156040895, 57227415], dtype=np.int64)   # chr1..22, X, Y
N_AUTOSOMES = 22
NUCLEUS_RADIUS = 5500.0    # demo/config_file.json:22
OCCUPANCY = 0.4            # demo/config_file.json:14


def genome_bins(resolution: int, scale: float = 1.0) -> np.ndarray:
    """Number of bins per chromosome (24 entries).  ``scale`` < 1 shrinks every
    chromosome proportionally (small test genomes with the same topology)."""
    nb = np.ceil(HG38_LENGTHS * scale / float(resolution)).astype(np.int64)
    return np.maximum(nb, 1)


def build_index(bins: np.ndarray, n_diploid_chroms: int = N_AUTOSOMES):
    """Returns (chrom_hap, chrom_bead, copy_bead, CopyIndex).

    Bead order as in the reference: all copy-0 chains (incl. X, Y) first, then
    the copy-1 autosome chains."""
    nchrom = len(bins)
    chrom_hap = np.repeat(np.arange(nchrom, dtype=np.int32), bins)
    n_hap = int(bins.sum())
    n_dip = int(bins[:n_diploid_chroms].sum())
    chrom_bead = np.concatenate([chrom_hap, chrom_hap[:n_dip]])
    copy_bead = np.concatenate([np.zeros(n_hap, np.int32), np.ones(n_dip, np.int32)])
    return chrom_hap, chrom_bead, copy_bead, CopyIndex.diploid(n_hap, n_dip)


def bead_radius(n_bead_total: int) -> np.float32:
    """Uniform radius from nuclear occupancy (igm/_preprocess.py:153-162):
    n * (4/3 pi r^3) = occupancy * (4/3 pi R^3)."""
    return np.float32(NUCLEUS_RADIUS * (OCCUPANCY / n_bead_total) ** (1.0 / 3.0))


def _fold(x: np.ndarray, half: float) -> np.ndarray:
    """Triangle-wave reflection of a walk into [-half, half]."""
    period = 4.0 * half
    y = np.mod(x + half, period)
    y = np.where(y > 2.0 * half, period - y, y)
    return y - half


def random_walk_coordinates(chrom_bead: np.ndarray, copy_bead: np.ndarray,
                            nstruct: int, radius: float, rng: np.random.Generator,
                            chunk: int = 256) -> np.ndarray:
    """(nbead, nstruct, 3) float32: each chromosome copy is a random walk with
    step ~ 2r from a random origin, reflected into the cube inscribed in the
    nucleus (|x| <= R/sqrt(3)), so cis distances are short and trans distances
    span 0..~11000 like the demo population."""
    nbead = len(chrom_bead)
    out = np.empty((nbead, nstruct, 3), dtype=np.float32)
    half = NUCLEUS_RADIUS / np.sqrt(3.0)
    key = chrom_bead.astype(np.int64) * 2 + copy_bead
    starts = np.concatenate([[0], np.nonzero(np.diff(key))[0] + 1, [nbead]])
    step = 2.0 * float(radius) / np.sqrt(3.0)
    for s0 in range(0, nstruct, chunk):
        s1 = min(nstruct, s0 + chunk)
        ns = s1 - s0
        for a, b in zip(starts[:-1], starts[1:]):
            origin = rng.uniform(-half, half, size=(1, ns, 3))
            steps = rng.standard_normal(size=(b - a, ns, 3)) * step
            steps[0] = 0.0
            walk = origin + np.cumsum(steps, axis=0)
            out[a:b, s0:s1] = _fold(walk, half).astype(np.float32)
    return out


def make_population(resolution: int = 200_000, nstruct: int = 1000, seed: int = 20261018,
                    genome_scale: float = 1.0, radius: Optional[float] = None) -> Population:
    bins = genome_bins(resolution, genome_scale)
    chrom_hap, chrom_bead, copy_bead, ci = build_index(bins)
    nbead = len(chrom_bead)
    r = bead_radius(nbead) if radius is None else np.float32(radius)
    rng = np.random.default_rng(seed)
    crd = random_walk_coordinates(chrom_bead, copy_bead, nstruct, float(r), rng)
    radii = np.full(nbead, r, dtype=np.float32)
    return Population(crd, radii, chrom_bead, ci, copy_bead)


def make_prob_matrix(chrom_hap: np.ndarray, seed: int = 20261018,
                     intra_c: float = 1.5, intra_alpha: float = 1.0,
                     intra_min: float = 0.008, inter_per_row: float = 520.0,
                     inter_lo: float = 1e-4, inter_hi: float = 0.05) -> ProbMatrix:
    """Synthetic .hcs-like strict-upper-triangle CSR, float32.

    intra: p_ij = min(1, c/|i-j|^alpha) stored while >= intra_min;
    inter: ``inter_per_row`` random trans partners per row on average,
    log-uniform in [inter_lo, inter_hi].
    """
    rng = np.random.default_rng(seed + 1)
    n = len(chrom_hap)
    dmax = int(np.floor((intra_c / intra_min) ** (1.0 / intra_alpha)))
    chrom_end = np.searchsorted(chrom_hap, chrom_hap, side="right")   # first bin of next chrom
    rows_l, cols_l, vals_l = [], [], []
    # intra band
    i = np.arange(n, dtype=np.int64)
    n_intra = np.minimum(chrom_end - i - 1, dmax)
    ri = np.repeat(i, n_intra)
    off = np.arange(n_intra.sum()) - np.repeat(np.cumsum(n_intra) - n_intra, n_intra) + 1
    rows_l.append(ri)
    cols_l.append(ri + off)
    vals_l.append(np.minimum(1.0, intra_c / off.astype(np.float64) ** intra_alpha))
    # inter: sample partners j > end of own chromosome
    n_after = n - chrom_end
    frac = inter_per_row * 2.0 / max(1, n)     # density so that mean count/row ~ inter_per_row
    cnt = rng.binomial(n_after, np.minimum(1.0, frac))
    rr = np.repeat(i, cnt)
    u = rng.random(cnt.sum())
    cc = np.repeat(chrom_end, cnt) + np.floor(u * np.repeat(n_after, cnt)).astype(np.int64)
    key = rr * n + cc
    key.sort()
    key = key[np.concatenate([[True], np.diff(key) != 0])]
    rr, cc = key // n, key % n
    pv = np.exp(rng.uniform(np.log(inter_lo), np.log(inter_hi), size=len(rr)))
    rows_l.append(rr)
    cols_l.append(cc)
    vals_l.append(pv)
    rows = np.concatenate(rows_l)
    cols = np.concatenate(cols_l)
    vals = np.concatenate(vals_l).astype(np.float32)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))]).astype(np.int64)
    return ProbMatrix(indptr, cols.astype(np.int32), vals, chrom_hap,
                      np.ones(n, dtype=np.float32))


def random_walk_coordinates_torch(chrom_bead: np.ndarray, copy_bead: np.ndarray, nstruct: int,
                                  radius: float, seed: int, device, chunk: int = 1024):
    """Same construction as random_walk_coordinates, generated on the GPU with
    torch (benchmark-sized populations: 29 838 x 10 000 x 3 floats take minutes
    on the host).  Different random stream than the NumPy generator; parity
    checks always compare against what is actually resident on the device."""
    import torch
    nbead = len(chrom_bead)
    half = NUCLEUS_RADIUS / float(np.sqrt(3.0))
    key = chrom_bead.astype(np.int64) * 2 + copy_bead
    first = np.concatenate([[True], np.diff(key) != 0])
    first_t = torch.from_numpy(first).to(device)
    seg_id = torch.from_numpy(np.cumsum(first) - 1).to(device)
    nseg = int(first.sum())
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    out = torch.empty((nbead, nstruct, 3), dtype=torch.float32, device=device)
    step = 2.0 * float(radius) / float(np.sqrt(3.0))
    period = 4.0 * half
    for s0 in range(0, nstruct, chunk):
        ns = min(chunk, nstruct - s0)
        steps = torch.randn((nbead, ns, 3), generator=gen, device=device, dtype=torch.float32) * step
        steps[first_t] = 0.0
        walk = torch.cumsum(steps, dim=0)
        # subtract the cumulative value at each segment start, add a random origin
        seg_start = walk[first_t]                       # (nseg, ns, 3)
        origin = (torch.rand((nseg, ns, 3), generator=gen, device=device) * 2.0 - 1.0) * half
        walk = walk - seg_start[seg_id] + origin[seg_id]
        y = torch.remainder(walk + half, period)
        y = torch.where(y > 2.0 * half, period - y, y) - half
        out[:, s0:s0 + ns] = y
        del steps, walk, y
    return out
