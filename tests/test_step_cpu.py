"""Host logic of the ActivationDistanceStep drop-in (no GPU): candidate filter,
plast lookup, file formats and naming, against restatements of the reference
lines (igm/steps/ActivationDistanceStep.py:111-194, 234-309)."""
import os

import numpy as np
import pytest
import scipy.sparse

from igm_b200 import hdf5, synthetic
from igm_b200.population import ProbMatrix
import importlib

S = importlib.import_module("igm_b200.steps.ActivationDistanceStep")   # the module (the package exports the class)
from igm_b200.steps._compat import Config
from oracle import actdist_oracle as orc
from tests import helpers as H


def _small_matrix(seed=0):
    bins = synthetic.genome_bins(2_000_000, 0.05)
    chrom_hap, _, _, _ = synthetic.build_index(bins)
    return synthetic.make_prob_matrix(chrom_hap, seed=seed, inter_per_row=20.0)


def test_filter_candidates_matches_reference_loop():
    pm = _small_matrix()
    # literal restatement of the setup loop, :166-178
    rows = pm.rows()
    keep_i, keep_j, keep_p = [], [], []
    intra_sigma, inter_sigma = 0.05, 0.02
    for i, j, pwish in zip(rows, pm.indices, pm.data):        # coo_generator order
        keep1 = (intra_sigma is not False) and (pm.chrom[i] == pm.chrom[j]) and (pwish >= intra_sigma)
        keep2 = (inter_sigma is not False) and (pm.chrom[i] != pm.chrom[j]) and (pwish >= inter_sigma)
        if keep1 or keep2:
            keep_i.append(i); keep_j.append(j); keep_p.append(pwish)
    ii, jj, pw = S.filter_candidates(pm, intra_sigma, inter_sigma)
    assert np.array_equal(ii, np.array(keep_i, np.int32))
    assert np.array_equal(jj, np.array(keep_j, np.int32))
    assert np.array_equal(pw, np.array(keep_p, np.float64))
    # and the oracle's vectorised form
    oi, oj, op = orc.select_candidates(pm.indptr, pm.indices, pm.data, pm.chrom, intra_sigma, inter_sigma)
    assert np.array_equal(ii, oi) and np.array_equal(jj, oj) and np.array_equal(pw, op)
    # one of the two lists switched off
    ii2, jj2, _ = S.filter_candidates(pm, 0.05, False)
    assert len(ii2) and np.all(pm.chrom[ii2] == pm.chrom[jj2])


@pytest.mark.skipif(not H.have_demo(), reason="oracle/_ref/demo not present")
def test_filter_on_demo_matrix_float32_edge():
    pm = ProbMatrix.from_hcs(H.DEMO_HCS)
    s = H.demo_full_summary()["sigmas"]
    for sig in ("1", "0.2", "0.1", "0.05", "0.02", "0.01"):
        ii, jj, pw = S.filter_candidates(pm, float(sig), float(sig))          # float32 compare
        assert len(ii) == s[sig]["pairs_f32_compare"]
        ii, jj, pw = S.filter_candidates(pm, float(sig), float(sig), np.float64)
        assert len(ii) == s[sig]["pairs_f64_compare"]


def test_lookup_plast_matches_coo_lil(tmp_path):
    rng = np.random.default_rng(3)
    n = 50
    # a previous actdist file: bead-index records, some with row/col >= n
    row = rng.integers(0, 2 * n, 400).astype(np.int32)
    col = rng.integers(0, 2 * n, 400).astype(np.int32)
    key = row.astype(np.int64) * 4 * n + col
    _, first = np.unique(key, return_index=True)
    row, col = row[first], col[first]
    prob = orc.text_roundtrip(rng.uniform(0, 1, len(row)))
    f = str(tmp_path / "actdist.hdf5")
    hdf5.write_h5(f, {"row": row, "col": col, "dist": np.zeros(len(row), np.float32), "prob": prob})
    # reference lines :145-156
    m = np.logical_and(row < n, col < n)
    plast = scipy.sparse.coo_matrix((prob[m], (row[m], col[m])), shape=(n, n)).tolil()
    ii = rng.integers(0, n, 300).astype(np.int32)
    jj = rng.integers(0, n, 300).astype(np.int32)
    exp = np.array([plast[a, b] for a, b in zip(ii, jj)], dtype=np.float64)
    got = S.lookup_plast(f, n, ii, jj)
    assert np.array_equal(got, exp)
    assert np.array_equal(S.lookup_plast(None, n, ii, jj), np.zeros(300))


@pytest.mark.parametrize("frac", [0.0, 0.3, 1.0])
def test_lookup_plast_sorted_fast_paths(tmp_path, frac):
    """Files this step writes list their records in candidate order (strictly increasing
    copy-0 keys) and usually hold fewer pairs than the next, smaller sigma selects: both
    shortcuts must give what the coo -> lil lookup of the reference gives (:145-156,177)."""
    rng = np.random.default_rng(int(frac * 10))
    n = 80
    ii, jj = np.triu_indices(n, 1)
    keep = np.sort(rng.choice(len(ii), 900, replace=False))
    ii, jj = ii[keep].astype(np.int32), jj[keep].astype(np.int32)        # sorted candidate list
    prev = np.sort(rng.choice(len(ii), int(frac * len(ii)), replace=False))
    # previous file: per pair one copy-0 record followed by a copy-1 record (row, col >= n)
    row = np.repeat(ii[prev], 2); col = np.repeat(jj[prev], 2)
    row[1::2] += n; col[1::2] += n
    prob = orc.text_roundtrip(rng.uniform(0, 1, len(row)))
    f = str(tmp_path / "actdist.hdf5")
    hdf5.write_h5(f, {"row": row.astype(np.int32), "col": col.astype(np.int32),
                      "dist": np.zeros(len(row), np.float32), "prob": prob})
    exp = np.zeros(len(ii))
    exp[prev] = prob[0::2].astype(np.float64)
    assert np.array_equal(S.lookup_plast(f, n, ii, jj), exp)
    # an unsorted candidate list takes the general path
    perm = rng.permutation(len(ii))
    assert np.array_equal(S.lookup_plast(f, n, ii[perm], jj[perm]), exp[perm])


def _cfg(tmp_path, hcs, hss, **hic):
    d = {"restraints": {"Hi-C": dict({"input_matrix": hcs, "intra_sigma_list": [0.2, 0.05],
                                      "inter_sigma_list": [0.2, 0.05], "contact_range": 2.0,
                                      "tmp_dir": "actdist", "keep_temporary_files": False}, **hic)},
         "optimization": {"structure_output": hss, "iter_corr_knob": 0},
         "parameters": {"workdir": str(tmp_path), "tmp_dir": str(tmp_path / "tmp")},
         "runtime": {"Hi-C": {}}}
    return Config(d)


def test_step_bookkeeping_name_setup_reduce_skip(tmp_path):
    pm = _small_matrix(1)
    hcs = str(tmp_path / "m.hcs")
    pm.save_hcs(hcs)
    cfg = _cfg(tmp_path, hcs, "unused.hss", gpu_shards=3, write_task_files=True)   # (a parallel controller's default)
    cfg["runtime"]["opt_iter"] = 2
    step = S.ActivationDistanceStep(cfg)
    # sigma lists are consumed exactly as the reference does (:69-98)
    assert cfg.get("runtime/Hi-C/inter_sigma") == 0.2 and cfg.get("runtime/Hi-C/inter_sigma_list") == [0.05]
    assert step.name() == "ActivationDistanceStep (INTER sigma=20.00%, INTRA sigma=20.00%, iter=2)"
    step.setup()
    assert list(step.argument_list) == [0, 1, 2]
    assert step.tmp_extensions == [".npy", ".tmp"]
    parts = [np.load(os.path.join(step.tmp_dir, "%d.in.npy" % b)) for b in range(3)]
    allp = np.concatenate(parts)
    ii, jj, pw = S.filter_candidates(pm, 0.2, 0.2)
    assert allp.dtype == S.PAIR_DTYPE and allp.shape == (len(ii),)     # one typed row per pair
    assert np.array_equal(allp["i"], ii) and np.array_equal(allp["pwish"], pw) and allp["j"].dtype == np.int32
    # ... or the reference's own float64 (n, 4) layout (:181-186) on request
    cfg_r = _cfg(tmp_path, hcs, "unused.hss", gpu_shards=3, reference_task_files=True, tmp_dir="actdist_ref")
    step_r = S.ActivationDistanceStep(cfg_r)
    step_r.setup()
    allr = np.concatenate([np.load(os.path.join(step_r.tmp_dir, "%d.in.npy" % b)) for b in range(3)])
    assert allr.dtype == np.float64 and allr.shape == (len(ii), 4)
    assert np.array_equal(allr[:, 0], ii) and np.array_equal(allr[:, 2], pw)
    # reduce: concatenation, dtypes, swap-file naming (:260-298)
    rng = np.random.default_rng(0)
    exp = []
    for b in range(3):
        rec = np.zeros(5 + b, dtype=S.actdist_shape)
        rec["row"] = rng.integers(0, 99, len(rec)); rec["col"] = rng.integers(0, 99, len(rec))
        rec["dist"] = rng.random(len(rec)); rec["prob"] = rng.random(len(rec))
        np.save(os.path.join(step.tmp_dir, "%d.out.npy" % b), rec)
        exp.append(rec)
    exp = np.concatenate(exp)
    old = os.path.join(step.tmp_dir, "actdist.hdf5")
    hdf5.write_h5(old, {"row": np.zeros(1, np.int32), "col": np.zeros(1, np.int32),
                        "dist": np.zeros(1, np.float32), "prob": np.zeros(1, np.float32)})
    cfg["runtime"]["Hi-C"]["actdist_file"] = old
    step.reduce()
    assert cfg["runtime"]["Hi-C"]["actdist_file"] == old
    assert os.path.exists(old + ".INTERsigma_0.2000.INTRAsigma_0.2000.iter_1")
    with hdf5.open_h5(old) as f:
        assert f["row"].dtype == np.int32 and f["dist"].dtype == np.float32
        for k in ("row", "col", "dist", "prob"):
            assert np.array_equal(f[k][()], exp[k])
    step.cleanup()
    assert not [x for x in os.listdir(step.tmp_dir) if x.endswith(".npy")]
    cfg2 = _cfg(tmp_path, hcs, "unused.hss")
    s2 = S.ActivationDistanceStep(cfg2)
    s2.skip()
    assert cfg2["runtime"]["Hi-C"]["actdist_file"] == old


def test_hdf5_writer_roundtrip_and_layout(tmp_path):
    f = str(tmp_path / "t.h5")
    a = {"row": np.arange(10, dtype=np.int32), "g/x": np.linspace(0, 1, 7).astype(np.float32),
         "g/y": np.arange(6, dtype=np.int64).reshape(2, 3), "empty": np.zeros(0, np.float32)}
    hdf5.write_h5(f, a, attrs={"nbead": np.int64(3), "version": np.int32(2)})
    with hdf5.open_h5(f) as h:
        assert set(h.keys()) == {"row", "g", "empty"}
        assert np.array_equal(h["row"][()], a["row"])
        assert np.array_equal(h["g"]["y"][()], a["g/y"]) and h["g"]["x"].dtype == np.float32
        assert h["empty"][()].shape == (0,)
        assert h.attrs["nbead"] == 3
    raw = open(f, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0     # classic superblock v0


def test_sprite_step_setup_and_reduce_host_logic(tmp_path):
    """SpriteAssignmentStep without a GPU: batch layout of setup, and reduce (Gibbs draw with
    occupancy penalty, igm/steps/SpriteAssignmentStep.py:166-259) on task files produced by
    the oracle-driven restatement - same seed, same assignment."""
    from igm_b200 import synthetic
    from igm_b200.steps import SpriteAssignmentStep
    pop = synthetic.make_population(2_000_000, 60, seed=5, genome_scale=0.02)
    hss = str(tmp_path / "p.hss")
    pop.save_hss(hss)
    rng = np.random.default_rng(9)
    chrom_hap, n_hap = pop.chrom_hap(), pop.n_hap
    clusters = []
    for k in range(11):
        cs = rng.choice(np.unique(chrom_hap), size=int(rng.integers(1, 6)), replace=False)
        pool = np.nonzero(np.isin(chrom_hap, cs))[0]
        clusters.append(np.sort(rng.choice(pool, size=int(min(len(pool), rng.integers(2, 9))), replace=False)).astype(np.int32))
    indptr = np.concatenate([[0], np.cumsum([len(c) for c in clusters])]).astype(np.int32)
    clf = str(tmp_path / "clusters.h5")
    hdf5.write_h5(clf, {"indptr": indptr, "data": np.concatenate(clusters).astype(np.int32)})
    cfg = Config({"parameters": {"workdir": str(tmp_path), "tmp_dir": str(tmp_path / "tmp")},
                           "optimization": {"structure_output": hss},
                           "restraints": {"sprite": {"clusters": clf, "volume_fraction_list": [0.5, 0.2],
                                                     "batch_size": 4, "keep_best": 10, "max_chrom_in_cluster": 3}},
                           "runtime": {"sprite": {}}})
    step = SpriteAssignmentStep(cfg)
    assert cfg.get("runtime/sprite/volume_fraction") == 0.5 and cfg.get("runtime/sprite/volume_fraction_list") == [0.2]
    assert step.name() == "SpriteAssignmentStep (volume_fraction=0.5%, iter=N/A)"
    step.setup()
    assert list(step.argument_list) == [0, 1, 2] and step.n_clusters == 11 and step.n_struct == 60
    assert step.tmp_extensions == [".npy", ".npz"]
    cidict = {i: pop.copy_index[i] for i in range(n_hap)}
    np.random.seed(31)
    batches = [H.sprite_reference_task(pop.coordinates, pop.chrom, cidict, clusters[b * 4:(b + 1) * 4], 10, 3)
               for b in range(3)]
    for b, (sel, idx, val) in enumerate(batches):          # the files task() writes (:155-161)
        np.savez(os.path.join(step.tmp_dir, 'tmp.%d.selected.npz' % b), *sel)
        np.save(os.path.join(step.tmp_dir, 'tmp.%d.idx.npy' % b), idx)
        np.save(os.path.join(step.tmp_dir, 'tmp.%d.values.npy' % b), val)
    np.random.seed(77)
    step.reduce()
    np.random.seed(77)
    exp_assign, exp_sel = H.sprite_reference_reduce(batches, indptr, 60, 4, 100.0)
    with hdf5.open_h5(os.path.join(step.tmp_dir, "assignment.h5")) as f:
        assert np.array_equal(np.asarray(f["assignment"][()]), exp_assign)
        assert np.array_equal(np.asarray(f["selected"][()]), exp_sel)
    assert (exp_assign >= 0).any()


@pytest.mark.parametrize("sig", [(0.2, 0.2), (0.05, 0.3), (0.3, 0.05), (False, 0.1), (0.1, False), (None, None)])
def test_native_filter_equals_numpy_form(sig):
    """igmk_filter_candidates (C) against the vectorised NumPy form, every sigma combination."""
    pm = _small_matrix(seed=4)
    a = S.filter_candidates(pm, sig[0], sig[1], native=True)
    b = S.filter_candidates(pm, sig[0], sig[1], native=False)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    if sig[0] and sig[1]:
        assert len(a[0]) > 0


def test_native_plast_join_equals_general_path(tmp_path):
    rng = np.random.default_rng(1)
    n = 120
    ii, jj = np.triu_indices(n, 1)
    keep = np.sort(rng.choice(len(ii), 2500, replace=False))
    ii, jj = ii[keep].astype(np.int32), jj[keep].astype(np.int32)
    prev = np.sort(rng.choice(len(ii), 1400, replace=False))
    row = np.repeat(ii[prev], 2); col = np.repeat(jj[prev], 2)
    row[1::2] += n
    prob = orc.text_roundtrip(rng.uniform(0, 1, len(row)))
    f = str(tmp_path / "actdist.hdf5")
    hdf5.write_h5(f, {"row": row.astype(np.int32), "col": col.astype(np.int32),
                      "dist": np.zeros(len(row), np.float32), "prob": prob})
    a = S.lookup_plast(f, n, ii, jj, native=True)
    b = S.lookup_plast(f, n, ii, jj, native=False)
    assert np.array_equal(a, b) and (a > 0).sum() > 1000
    perm = rng.permutation(len(ii))                      # unsorted candidates: the C join declines
    assert np.array_equal(S.lookup_plast(f, n, ii[perm], jj[perm], native=True), b[perm])


def test_hdf5_writer_roundtrip_property(tmp_path):
    """The streaming writer against the reader over random dataset sets: dtypes, shapes
    (incl. empty and 2-D), one group level, attributes; every dataset is 8-byte aligned."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    dtypes = [np.int32, np.int64, np.float32, np.float64, np.uint8, np.int16]

    @settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.lists(st.tuples(st.sampled_from(range(len(dtypes))), st.integers(0, 3000), st.integers(1, 3),
                              st.booleans()), min_size=1, max_size=7),
           st.integers(0, 2 ** 31 - 1))
    def run(spec, seed):
        rng = np.random.default_rng(seed)
        data = {}
        for k, (di, n, cols, grouped) in enumerate(spec):
            shape = (n,) if cols == 1 else (n, cols)
            a = (rng.integers(-1000, 1000, size=shape)).astype(dtypes[di])
            data[("g/" if grouped else "") + "d%d" % k] = a
        f = str(tmp_path / "p.h5")
        hdf5.write_h5(f, data, attrs={"n": np.int64(len(data))})
        with hdf5.open_h5(f) as h:
            assert h.attrs["n"] == len(data)
            for name, a in data.items():
                node = h
                for part in name.split("/"):
                    node = node[part]
                got = node[()]
                assert got.dtype == a.dtype and got.shape == a.shape and np.array_equal(got, a)
                if a.size:
                    assert np.array_equal(node[:max(1, a.shape[0] // 2)], a[:max(1, a.shape[0] // 2)])
    run()


def test_engine_cache_keeps_one_engine_per_device(tmp_path, monkeypatch):
    """_get_engine: shards of one A-step on different devices keep their staged populations;
    a rewritten or different .hss replaces all of them."""
    made, closed, moved = [], [], []

    class FakeEngine:
        def __init__(self, path, device):
            self.tag = (path, device)
            self._pending_xyz = None
            made.append(self.tag)

        def upload_coordinates(self, xyz, bead0=0):
            moved.append(("h2d", self.tag[1], bead0, len(xyz)))

        def copy_coordinates_from(self, other, bead0, nb):
            moved.append(("p2p", self.tag[1], other.tag[1], bead0, nb))

        def close(self):
            closed.append(self.tag)

    monkeypatch.setattr(S, "ActdistEngine",
                        type("E", (), {"from_hss": staticmethod(lambda p, d, staged=None, upload=True: FakeEngine(p, d))}))
    monkeypatch.setattr(S, "_engine_cache", {})
    monkeypatch.setattr(S, "_staged_cache", {})
    staged = []
    import types
    monkeypatch.setattr(S, "_stage_population", lambda p: staged.append(p) or types.SimpleNamespace(
        nbead=11, coordinates=np.zeros((11, 2, 3), np.float32)))
    a, b = str(tmp_path / "a.hss"), str(tmp_path / "b.hss")
    open(a, "wb").write(b"x" * 10)
    open(b, "wb").write(b"y" * 20)
    e0 = S._get_engine(a, 0)
    e1 = S._get_engine(a, 1)
    assert S._get_engine(a, 0) is e0 and S._get_engine(a, 1) is e1 and not closed
    assert staged == [a]                                  # one host copy of the file for both devices
    e23 = S._get_engines(a, [2, 3])                       # staged concurrently, same host copy
    assert [e.tag for e in e23] == [(a, 2), (a, 3)] and staged == [a]
    # each new device uploads its share of the beads and pulls the other share from its peer
    assert sorted(moved) == [("h2d", 2, 0, 6), ("h2d", 3, 6, 5), ("p2p", 2, 3, 6, 5), ("p2p", 3, 2, 0, 6)]
    S._get_engine(b, 0)                                   # another population: all are released
    assert sorted(closed) == [(a, 0), (a, 1), (a, 2), (a, 3)] and len(S._engine_cache) == 1
    assert staged == [a, b] and len(S._staged_cache) == 1


def test_dropin_runs_under_the_reference_step_run(tmp_path):
    """The drop-in subclasses the reference's REAL igm.core.Step and is driven by its
    Step.run (igm/core/step.py:226-322): sqlite restart log (job_tracking.py), statuses in the
    reference's order, and a second run from a fresh configuration takes skip() and restores
    runtime/Hi-C/actdist_file without calling task again.  Here the device call is replaced
    by the NumPy oracle (no GPU); tests/test_gpu_step.py runs the same driver on the kernels."""
    import json
    import subprocess
    import sys
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("no copy of the reference package (run `make -C oracle`)")
    r = subprocess.run([sys.executable, os.path.join(H.ROOT, "tests", "real_step_driver.py"), str(tmp_path), "--fake-gpu"],
                       capture_output=True, text=True, timeout=600, cwd=H.ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["records"] > 0 and d["records_equal_oracle"]
    assert d["statuses"] == ["entry", "setup", "map", "mapped", "reduced", "cleanup", "completed"]
    assert d["task_calls_first_run"] == 1 and d["task_calls_second_run"] == 0
    assert d["second_run_actdist_file"] == d["actdist_file"] and os.path.exists(d["actdist_file"])
    assert d["second_run_sigma"] == 0.05


def test_engine_cache_deferred_upload(tmp_path, monkeypatch):
    """_get_engines(defer_upload=True) creates engines whose coordinates stay in host memory
    (the A-step's own device call stages them); any other user of the cache gets them staged."""
    log = []

    class FakeEngine:
        def __init__(self, path, device, upload):
            self._pending_xyz = None if upload else "xyz"
            log.append(("new", device, upload))

        def stage_pending(self):
            log.append(("stage", self._pending_xyz))
            self._pending_xyz = None

        def close(self):
            pass

    monkeypatch.setattr(S, "ActdistEngine", type("E", (), {
        "from_hss": staticmethod(lambda p, d, staged=None, upload=True: FakeEngine(p, d, upload))}))
    monkeypatch.setattr(S, "_engine_cache", {})
    monkeypatch.setattr(S, "_staged_cache", {})
    monkeypatch.setattr(S, "_stage_population", lambda p: object())
    a = str(tmp_path / "a.hss")
    open(a, "wb").write(b"x" * 10)
    e = S._get_engines(a, [0], defer_upload=True)[0]
    assert e._pending_xyz == "xyz" and log == [("new", 0, False)]
    assert S._get_engines(a, [0], defer_upload=True)[0] is e and e._pending_xyz == "xyz"
    assert S._get_engine(a, 0) is e and e._pending_xyz is None      # e.g. the SPRITE step: staged now
    assert log[-1] == ("stage", "xyz")
    S._get_engines(a, [1])
    assert log[-1] == ("new", 1, True)


def test_threaded_host_phases_match_numpy_forms(tmp_path):
    """Lists long enough for the C library to split the candidate filter and the plast join over
    several host threads (> 65 536 entries per thread): same output as the NumPy forms."""
    from igm_b200 import hdf5
    rng = np.random.default_rng(5)
    n, per = 3000, 170
    cols = np.sort(np.stack([rng.choice(n, per, replace=False) for _ in range(n)]), axis=1)
    indptr = np.arange(0, n * per + 1, per, dtype=np.int64)
    data = rng.uniform(0, 0.05, n * per).astype(np.float32)
    chrom = np.sort(rng.integers(0, 23, n)).astype(np.int32)
    pm = S.ProbMatrix(indptr, cols.reshape(-1).astype(np.int32), data, chrom)
    a = S.filter_candidates(pm, 0.01, 0.02, native=True)
    b = S.filter_candidates(pm, 0.01, 0.02, native=False)
    assert len(a[0]) > 200000
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    # previous-iteration file: every third candidate has a stored probability, plus records of
    # second copies (row / col >= n) in between, which the join must skip
    ii, jj, _ = a
    keep = np.arange(len(ii)) % 3 == 0
    row = np.repeat(ii[keep], 2)
    col = np.repeat(jj[keep], 2)
    row[1::2] += n
    col[1::2] += n
    prob = rng.uniform(0, 1, len(row)).astype(np.float32)
    f = str(tmp_path / "prev.hdf5")
    hdf5.write_h5(f, {"row": row.astype(np.int32), "col": col.astype(np.int32),
                      "dist": np.zeros(len(row), np.float32), "prob": prob})
    p1 = S.lookup_plast(f, n, ii, jj, native=True)
    p0 = S.lookup_plast(f, n, ii, jj, native=False)
    assert np.array_equal(p1, p0) and np.count_nonzero(p1) > 60000
    # an unsorted candidate list makes the native join decline (the general path answers)
    perm = rng.permutation(len(ii))
    assert np.array_equal(S.lookup_plast(f, n, ii[perm], jj[perm], native=True), p0[perm])
