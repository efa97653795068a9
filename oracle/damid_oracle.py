"""TEST INFRASTRUCTURE ONLY - CPU restatement (NumPy) of the lamina DamID
activation distance for a spherical envelope,
igm/steps/DamidActivationDistanceStep.py:375-470 (``get_damid_actdist_I``),
``snormsq_sphere`` :39-55, ``cleanProbability`` :362-372, and of the text round
trip of ``task`` / ``reduce`` (:30-35, :266-270, :287-296).

Pinned: tests/golden/make_golden_damid.py runs the reference's own
``get_damid_actdist_I`` (imported through oracle/ref_loader.py) and stores its
outputs in tests/golden/damid_small.npz; tests/test_oracle_golden.py compares.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline may
import this module; the product (igm_b200) never does.

Arithmetic notes (NumPy >= 2, NEP 50), all reproduced on the device:
  * the batch file is float32 (:178-186), so I, p_exp and plast reach the function
    as np.float32 scalars; ``pnow`` is a Python float, hence
    ``(pnow - plast) / (1.0 - plast)`` is evaluated in float32, as is
    ``n_copies * n_struct * p``; ``round`` is half-to-even;
  * ``np.sum(np.square(x), axis=1)`` is the sequential float32 sum
    ((x^2 + y^2) + z^2); dividing by the 0-d float64 array ``(R - r)**2``
    promotes to float64; the order statistic is taken in DESCENDING order;
  * ellipsoids never reach this function in the reference (the test at :245 reads
    ``'shape' == 'ellipsoid'``, a comparison of two string literals).
"""
import numpy as np

DAMID_FMT_STR = "%6d %.5f %.5f"       # DamidActivationDistanceStep.py:35


def clean_probability(pij, pexist):    # :362-372
    if pexist < 1:
        pclean = (pij - pexist) / (1.0 - pexist)
    else:
        pclean = pij
    return max(0, pclean)


def get_damid_actdist_detail(I, p_exp, plast, coords, radii, copy_index, it_corr,
                             contact_range=0.05, nucleus_radius=5000.0):
    """Returns (records, detail): records = [(i, actdist, p) for i in copies(I)];
    detail = dict(s_bits, contact_count, o, p, denom)."""
    n_struct = coords.shape[1]
    ii = copy_index[I]
    n_copies = len(ii)
    r = radii[ii[0]]
    d_sq = np.empty(n_copies * n_struct)
    s_all = np.empty(n_copies * n_struct, dtype=np.float32)
    denom = None
    for i in range(n_copies):
        x = coords[ii[i]]
        R = np.array(nucleus_radius) * (1 - contact_range)
        s = np.sum(np.square(x), axis=1)                     # float32, sequential
        denom = (R - r) ** 2                                  # 0-d float64
        d_sq[i * n_struct:(i + 1) * n_struct] = s / denom
        s_all[i * n_struct:(i + 1) * n_struct] = s
    rcutsq = 1.0
    order = np.argsort(-d_sq, kind="stable")
    d_sq[::-1].sort()                                        # descending
    contact_count = int(np.count_nonzero(d_sq >= rcutsq))
    if it_corr == 1:
        pnow = float(contact_count) / (n_struct * n_copies)
        t = clean_probability(pnow, plast)
        p = clean_probability(p_exp, t)
    else:
        p = p_exp
    activation_distance = 2
    o = -1
    s_bits = 0
    if p > 0:
        o = min(n_copies * n_struct - 1, int(round(n_copies * n_struct * p)))
        activation_distance = np.sqrt(d_sq[o])
        # the float32 sum of squares behind d_sq[o] (ties share the value)
        s_sel = np.sort(s_all)[::-1][o]
        assert float(np.float64(s_sel) / denom) == float(d_sq[o])
        s_bits = int(np.float32(s_sel).view(np.uint32))
    recs = [(int(i), activation_distance, p) for i in ii]
    return recs, dict(s_bits=s_bits, contact_count=contact_count, o=o, p=float(p),
                      denom=float(denom), ncopies=n_copies)


def run_loci(loci, p_exp, plast, coords, radii, copy_index, it_corr, contact_range=0.05,
             nucleus_radius=5000.0):
    """loci / p_exp / plast as the float32 batch rows of setup (:178-186)."""
    params = np.array(list(zip(loci, p_exp, plast)), dtype=np.float32).reshape(-1, 3)
    recs, dets = [], []
    for I, pe, pl in params:
        r, d = get_damid_actdist_detail(int(I), pe, pl, coords, radii, copy_index, it_corr,
                                        contact_range, nucleus_radius)
        recs += r
        dets.append(d)
    return recs, dets


def task_text(records) -> str:         # :266-270
    return "\n".join([DAMID_FMT_STR % x for x in records])


def text_roundtrip5(values) -> np.ndarray:
    """float32(float('%.5f' % v)): what reduce's genfromtxt stores (:287-296)."""
    return np.array([float("%.5f" % v) for v in np.asarray(values, dtype=np.float64).ravel()],
                    dtype=np.float32)
