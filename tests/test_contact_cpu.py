"""Host-side logic of the contact-frequency path (no GPU): row scheduling across
ranks, count -> .hcs layout, evaluation statistics, oracle copy projection."""
import numpy as np

from igm_b200.contact import counts_to_probmatrix, evaluation_stats, row_blocks
from igm_b200.population import ProbMatrix
from oracle import contact_oracle as co


def test_row_blocks_cover_and_balance():
    n, block = 15453, 512
    for world in (1, 2, 4, 8):
        seen = np.zeros(n, int)
        work = []
        for r in range(world):
            w = 0
            for r0, r1 in row_blocks(n, block, r, world):
                seen[r0:r1] += 1
                w += (r1 - r0) * (n - r0)          # upper-triangle columns of the block row
            work.append(w)
        assert np.all(seen == 1)
        assert max(work) / (sum(work) / world) < 1.05    # boustrophedon pairing balances the triangle


def test_counts_to_probmatrix_layout(tmp_path):
    rng = np.random.default_rng(0)
    n = 23
    c = rng.integers(0, 50, (n, n)).astype(np.uint32)
    c = np.triu(c) + np.triu(c, 1).T
    c[rng.random((n, n)) < 0.3] = 0
    c = np.triu(c) + np.triu(c, 1).T
    chrom = np.repeat([0, 1, 2], [10, 8, 5]).astype(np.int32)
    pm = counts_to_probmatrix(c, 40, chrom, clip=True)
    assert pm.indptr[-1] == len(pm.indices) == np.count_nonzero(np.triu(c, 1))
    assert np.all(pm.indices > pm.rows())                 # strict upper triangle, CSR order
    dense = np.zeros((n, n), np.float32)
    dense[pm.rows(), pm.indices] = pm.data
    assert np.array_equal(dense, np.triu((c / 40.0).clip(0, 1), 1).astype(np.float32))
    p = str(tmp_path / "m.hcs")
    pm.save_hcs(p)
    back = ProbMatrix.from_hcs(p)
    assert np.array_equal(back.data, pm.data) and np.array_equal(back.indices, pm.indices)
    assert np.array_equal(back.indptr, pm.indptr)


def test_evaluation_stats_matches_loop():
    rng = np.random.default_rng(1)
    n = 30

    def rand_pm(density):
        m = np.triu(rng.random((n, n)) < density, 1)
        i, j = np.nonzero(m)
        indptr = np.zeros(n + 1, np.int64)
        np.cumsum(np.bincount(i, minlength=n), out=indptr[1:])
        return ProbMatrix(indptr, j.astype(np.int32), rng.uniform(0.001, 1, len(i)).astype(np.float32),
                          np.zeros(n, np.int32))
    inp, out = rand_pm(0.5), rand_pm(0.6)
    sigma = 0.2
    # the reference's loop (igm/steps/HicEvaluationStep.py:156-166)
    d = {(i, j): p for i, j, p in zip(inp.rows(), inp.indices, inp.data) if p >= np.float32(sigma) and i != j}
    diffs, rel = [], []
    for i, j, po in zip(out.rows(), out.indices, out.data):
        p = d.get((i, j))
        if p is not None:
            diffs.append(float(po) - float(p))
            rel.append((float(po) - float(p)) / float(p))
    score, avg, avg_rel = evaluation_stats(inp, out, sigma)
    assert np.isclose(score, np.abs(rel).mean()) and np.isclose(avg, np.mean(diffs)) and np.isclose(avg_rel, np.mean(rel))


def test_oracle_sum_copies_small():
    counts = np.arange(16, dtype=np.uint32).reshape(4, 4)
    ptr, beads = np.array([0, 2, 3, 4]), np.array([0, 3, 1, 2])       # locus 0 = beads {0, 3}
    hap = co.sum_copies(counts, ptr, beads)
    assert hap[0, 0] == counts[0, 0] + counts[0, 3] + counts[3, 0] + counts[3, 3]
    assert hap[0, 1] == counts[0, 1] + counts[3, 1] and hap[2, 1] == counts[2, 1]


def test_hss_roundtrip(tmp_path):
    from igm_b200 import synthetic
    from igm_b200.population import Population
    pop = synthetic.make_population(2_000_000, 9, seed=4, genome_scale=0.012)
    p = str(tmp_path / "pop.hss")
    pop.save_hss(p)
    q = Population.from_hss(p)
    assert np.array_equal(q.coordinates, pop.coordinates) and np.array_equal(q.radii, pop.radii)
    assert q.copy_index.to_dict() == pop.copy_index.to_dict()
    assert np.array_equal(q.chrom_hap(), pop.chrom_hap())
