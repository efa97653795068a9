"""TEST INFRASTRUCTURE ONLY - CPU restatement (NumPy) of the reference A-step.

This file is the checker for the CUDA path; it is never imported by the
product package ``igm_b200`` (only by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs).

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the reference's
own, unmodified ``get_actdist`` functions (imported from /root/reference via
``oracle/ref_loader.py``) on the shipped demo population and on seeded
synthetic populations; ``tests/test_oracle_golden.py`` checks this restatement
against those committed vectors, and ``tests/test_oracle_vs_reference.py``
compares it with the live reference whenever /root/reference is present.

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import hashlib
from typing import List, Sequence, Tuple

import numpy as np

# igm/steps/ActivationDistanceStep.py:38
ACTDIST_FMT_STR = "%6d %6d %10.4f %.4f"
# igm/steps/GP_activation.py:33
ACTDIST_FMT_STR_GP = "%6d %6d %10.2f %.5f"
# igm/steps/ActivationDistanceStep.py:32-37
ACTDIST_SHAPE = [("row", "int32"), ("col", "int32"), ("dist", "float32"), ("prob", "float32")]

MODE_LB = 0   # igm/steps/ActivationDistanceStep.py:336-485 (active in igm-run)
MODE_GP = 1   # igm/steps/GP_activation.py:317-445, igm/utils/actdist.py:15-96


def clean_probability(pij, pexist):
    """igm/steps/ActivationDistanceStep.py:314-332."""
    if pexist < 1:
        pclean = (pij - pexist) / (1.0 - pexist)
    else:
        pclean = pij
    return max(0, pclean)


class PairDetail:
    """Everything the CUDA kernel has to reproduce for one candidate pair."""
    __slots__ = ("records", "d2_sel_bits", "contact_count", "o", "p", "npc", "nrec")

    def __init__(self):
        self.records: List[Tuple[int, int, float, float]] = []
        self.d2_sel_bits = 0
        self.contact_count = 0
        self.o = -1
        self.p = 0.0
        self.npc = 0
        self.nrec = 0


def get_actdist_detail(i, j, pwish, plast, coords, radii, chrom_hap, copy_index,
                       it_corr, contact_range=2, mode=MODE_LB) -> PairDetail:
    """One candidate pair.

    LB mode follows igm/steps/ActivationDistanceStep.py:379-485;
    GP mode follows igm/steps/GP_activation.py:355-445
    (== igm/utils/actdist.py:46-96 when it_corr == 1).

    coords: (nbead, nstruct, 3) float32; copy_index[i] -> list of bead ids.
    """
    det = PairDetail()
    if i == j:                                           # :379-380
        return det
    n_struct = coords.shape[1]                           # :382
    ii = copy_index[i]                                   # :389-390
    jj = copy_index[j]
    ri, rj = radii[ii[0]], radii[jj[0]]                  # :392-393
    rcutsq = np.square(contact_range * (ri + rj))        # :396 (float32 under NumPy >= 2)
    intra = chrom_hap[i] == chrom_hap[j]                 # :405

    if mode == MODE_LB and intra:
        # :405-419 - only (i,j) and (i',j')
        n_combinations = len(ii)
        n_possible_contacts = min(len(ii), len(jj))
        if len(ii) != len(jj):
            # the reference would read uninitialised memory here.
            raise ValueError("LB intra pair with unequal copy counts (%d,%d)" % (i, j))
        d_sq = np.empty((n_combinations, n_struct))
        it = 0
        for k, m in zip(ii, jj):
            x = coords[k]
            y = coords[m]
            d_sq[it] = np.sum(np.square(x - y), axis=1)
            it += 1
    else:
        # LB inter :421-436  /  GP :364-383 (all combinations)
        n_combinations = len(ii) * len(jj)
        if mode == MODE_LB:
            n_possible_contacts = len(ii) * len(jj)      # :424
        else:
            n_possible_contacts = min(len(ii), len(jj))  # GP_activation.py:366
        d_sq = np.empty((n_combinations, n_struct))
        it = 0
        for k in ii:
            for m in jj:
                x = coords[k]
                y = coords[m]
                d_sq[it] = np.sum(np.square(x - y), axis=1)
                it += 1

    d_sq.sort(axis=0)                                    # :439
    contact_count = np.count_nonzero(d_sq[0:n_possible_contacts, :] <= rcutsq)   # :442
    pnow = float(contact_count) / (n_possible_contacts * n_struct)               # :445
    sortdist_sq = np.sort(d_sq[0:n_possible_contacts, :].ravel())                # :448

    if it_corr == 1:                                     # :452-462
        t = clean_probability(pnow, plast)
        p = clean_probability(pwish, t)
    else:
        p = pwish

    det.contact_count = int(contact_count)
    det.npc = n_possible_contacts
    det.p = float(p)
    if p > 0:                                            # :466
        o = min(n_possible_contacts * n_struct - 1,
                int(round(n_possible_contacts * p * n_struct)))                  # :469-470
        activation_distance = np.sqrt(sortdist_sq[o])    # :473 (float64 sqrt)
        det.o = o
        det.d2_sel_bits = int(np.float32(sortdist_sq[o]).view(np.uint32))
        if intra:                                        # :476-478 (option == 0)
            det.records = [(i0, i1, activation_distance, p) for i0, i1 in zip(ii, jj)]
        else:                                            # :483
            det.records = [(i0, i1, activation_distance, p) for i0 in ii for i1 in jj]
    det.nrec = len(det.records)
    return det


def get_actdist(i, j, pwish, plast, coords, radii, chrom_hap, copy_index,
                it_corr, contact_range=2, mode=MODE_LB):
    """Same return value as the reference's get_actdist: list of (i, j, ad, p)."""
    return get_actdist_detail(i, j, pwish, plast, coords, radii, chrom_hap,
                              copy_index, it_corr, contact_range, mode).records


def task_text(records: Sequence[Tuple[int, int, float, float]], fmt=ACTDIST_FMT_STR) -> str:
    """igm/steps/ActivationDistanceStep.py:228-230: the '%d.out.tmp' payload."""
    return "\n".join([fmt % x for x in records])


def text_sha(records, fmt=ACTDIST_FMT_STR) -> str:
    return hashlib.sha256(task_text(records, fmt).encode()).hexdigest()[:16]


def text_roundtrip(values, ndec: int = 4) -> np.ndarray:
    """What reduce() stores: value -> '%.{ndec}f' text -> genfromtxt float32
    (igm/steps/ActivationDistanceStep.py:230,249).  Scalar formatting loop -
    slow but definitional."""
    fmt = "%." + str(ndec) + "f"
    return np.array([float(fmt % v) for v in values], dtype=np.float64).astype(np.float32)


def select_candidates(indptr, indices, data, chrom, intra_sigma, inter_sigma,
                      compare_dtype="float32"):
    """Candidate filter of setup(), igm/steps/ActivationDistanceStep.py:171-178.

    ``coo_generator`` order = CSR row-major.  ``pwish >= sigma`` is evaluated on
    the scalar type the generator yields; with float32 data under NumPy >= 2 a
    Python-float sigma is cast to float32 ('filter edge'),
    hence ``compare_dtype``.  ``i == j`` entries are dropped (quirk q6).
    Returns (i, j, pwish_float64) arrays.
    """
    n = len(indptr) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    cols = np.asarray(indices, dtype=np.int64)
    dt = np.dtype(compare_dtype)
    pw = np.asarray(data).astype(dt)
    intra = chrom[rows] == chrom[cols]
    keep = np.zeros(len(rows), dtype=bool)
    if intra_sigma is not False and intra_sigma is not None:
        keep |= intra & (pw >= dt.type(intra_sigma))
    if inter_sigma is not False and inter_sigma is not None:
        keep |= (~intra) & (pw >= dt.type(inter_sigma))
    keep &= rows != cols
    return (rows[keep].astype(np.int32), cols[keep].astype(np.int32),
            np.asarray(data)[keep].astype(np.float64))


def run_pairs(ii, jj, pwish, plast, coords, radii, chrom_hap, copy_index,
              it_corr, contact_range=2.0, mode=MODE_LB):
    """Loop of task(), igm/steps/ActivationDistanceStep.py:215-222, returning
    the flat record list plus per-pair details."""
    records = []
    details = []
    for a, b, pw, pl in zip(ii, jj, pwish, plast):
        det = get_actdist_detail(int(a), int(b), np.float64(pw), np.float64(pl), coords,
                                 radii, chrom_hap, copy_index, it_corr,
                                 contact_range, mode)
        details.append(det)
        records.extend(det.records)
    return records, details


def details_to_arrays(details):
    n = len(details)
    out = {
        "d2_sel_bits": np.zeros(n, np.uint32),
        "contact_count": np.zeros(n, np.int32),
        "o": np.full(n, -1, np.int32),
        "p": np.zeros(n, np.float64),
        "nrec": np.zeros(n, np.int32),
    }
    for k, d in enumerate(details):
        out["d2_sel_bits"][k] = d.d2_sel_bits
        out["contact_count"][k] = d.contact_count
        out["o"][k] = d.o
        out["p"][k] = d.p
        out["nrec"][k] = d.nrec
    return out


def records_to_arrays(records):
    """reduce(), igm/steps/ActivationDistanceStep.py:249-257: the four columns
    as stored (after the 4-decimal text round trip)."""
    if not records:
        return (np.zeros(0, np.int32), np.zeros(0, np.int32),
                np.zeros(0, np.float32), np.zeros(0, np.float32))
    row = np.array([r[0] for r in records], dtype=np.int32)
    col = np.array([r[1] for r in records], dtype=np.int32)
    dist = text_roundtrip([r[2] for r in records])
    prob = text_roundtrip([r[3] for r in records])
    return row, col, dist, prob
