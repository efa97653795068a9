"""End-to-end run of the Step drop-in on the GPU: setup -> task -> reduce over the
shipped demo population (config 1), compared with the reference's records."""
import hashlib
import os
import shutil

import numpy as np
import pytest

from oracle import actdist_oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not H.have_demo(), reason="oracle/_ref/demo not present")
def test_step_run_demo_two_iterations(tmp_path):
    from igm_b200 import hdf5
    from igm_b200.population import Population, ProbMatrix
    from igm_b200.steps.ActivationDistanceStep import ActivationDistanceStep, filter_candidates
    from igm_b200.steps._compat import Config
    hss = str(tmp_path / "igm-model.hss")
    shutil.copyfile(H.DEMO_HSS, hss)
    cfg = Config({
        "restraints": {"Hi-C": {"input_matrix": H.DEMO_HCS, "intra_sigma_list": [0.2, 0.1],
                                "inter_sigma_list": [0.2, 0.1], "contact_range": 2.0,
                                "tmp_dir": "actdist", "keep_temporary_files": True,
                                "gpu_shards": 2, "write_text_tmp": True}},
        "optimization": {"structure_output": hss, "iter_corr_knob": 1},
        "parameters": {"workdir": str(tmp_path), "tmp_dir": str(tmp_path / "tmp")},
        "runtime": {"Hi-C": {}}})
    pop = Population.from_hss(hss)
    pm = ProbMatrix.from_hcs(H.DEMO_HCS)
    plast_file = None
    for sigma in (0.2, 0.1):
        step = ActivationDistanceStep(cfg)
        assert cfg.get("runtime/Hi-C/inter_sigma") == sigma
        step.run()
        out = cfg["runtime"]["Hi-C"]["actdist_file"]
        # oracle for the same iteration (float32 sigma compare, it_corr = 1, plast from the previous file)
        from igm_b200.steps.ActivationDistanceStep import lookup_plast
        ii, jj, pw = filter_candidates(pm, sigma, sigma)
        pl = lookup_plast(plast_file, pm.n, ii, jj) if plast_file else np.zeros(len(ii))
        recs, _ = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(),
                                pop.copy_index, 1, 2.0, orc.MODE_LB)
        row, col, dist, prob = orc.records_to_arrays(recs)
        with hdf5.open_h5(out) as f:
            assert np.array_equal(f["row"][()], row) and np.array_equal(f["col"][()], col)
            assert np.array_equal(f["dist"][()].view(np.uint32), dist.view(np.uint32))
            assert np.array_equal(f["prob"][()].view(np.uint32), prob.view(np.uint32))
        # the reference's text wire format, byte for byte (two shards concatenated)
        text = "\n".join(open(os.path.join(step.tmp_dir, "%d.out.tmp" % b)).read() for b in range(2))
        assert hashlib.sha256(text.encode()).hexdigest() == hashlib.sha256(orc.task_text(recs).encode()).hexdigest()
        # genfromtxt on that text gives the stored columns (what the reference's reduce would store)
        g = np.genfromtxt(os.path.join(step.tmp_dir, "0.out.tmp"), dtype=orc.ACTDIST_SHAPE)
        with hdf5.open_h5(out) as f:
            assert np.array_equal(g["dist"], f["dist"][()][:len(g)])
            assert np.array_equal(g["prob"], f["prob"][()][:len(g)])
        # keep a copy as "previous iteration" (reduce moves the old file to the swap name)
        plast_file = str(tmp_path / ("prev_%g.hdf5" % sigma))
        shutil.copyfile(out, plast_file)
        # next A-step: the runtime sigma is consumed by the driver (bin/igm-run:175-305)
        del cfg["runtime"]["Hi-C"]["inter_sigma"], cfg["runtime"]["Hi-C"]["intra_sigma"]
        cfg["runtime"].pop("current_iteration_name", None)


def test_engine_from_hss_streams_chunks(tmp_path):
    """ActdistEngine.from_hss (chunk-wise staging, no host copy of the population) gives
    the same answers as the Population path; iter_chunks covers the dataset."""
    import numpy as np
    from igm_b200 import hdf5, synthetic
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, 90, seed=6, genome_scale=0.02)
    p = str(tmp_path / "pop.hss")
    pop.save_hss(p)
    with hdf5.open_h5(p) as f:
        got = np.zeros_like(pop.coordinates)
        for offs, ch in f["coordinates"].iter_chunks():
            got[offs[0]:offs[0] + ch.shape[0]] = ch
        assert np.array_equal(got, pop.coordinates)
    rng = np.random.default_rng(0)
    ii = rng.integers(0, pop.n_hap - 1, 300).astype(np.int32)
    jj = (ii + 1 + rng.integers(0, 5, 300)).clip(max=pop.n_hap - 1).astype(np.int32)
    nc, ch = pop.copy_index.ncopies(), pop.chrom_hap()
    ok = ~((ch[ii] == ch[jj]) & (nc[ii] != nc[jj])) & (ii != jj)
    ii, jj = ii[ok], jj[ok]
    pw = rng.uniform(0.01, 1, len(ii))
    with ActdistEngine(pop, 0) as a, ActdistEngine.from_hss(p, 0) as b:
        assert a.actdist(ii, jj, pw).tobytes() == b.actdist(ii, jj, pw).tobytes()
        assert np.array_equal(a.contact_counts_haploid(0, 9, 0, 9), b.contact_counts_haploid(0, 9, 0, 9))


def test_task_drives_all_gpus_concurrently(tmp_path):
    """One task, several devices (one host thread and one staged engine per device, the
    population read once into pinned memory): the output file equals the single-device one.
    Needs >= 2 GPUs on the box; skipped otherwise."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from igm_b200 import hdf5, synthetic
    from igm_b200.steps.ActivationDistanceStep import ActivationDistanceStep
    from igm_b200.steps._compat import Config
    pop = synthetic.make_population(1_000_000, 400, seed=12, genome_scale=0.1)
    hss = str(tmp_path / "pop.hss")
    pop.save_hss(hss)
    pm = synthetic.make_prob_matrix(pop.chrom_hap(), seed=3, inter_per_row=40.0, intra_min=0.002, inter_lo=0.005)
    hcs = str(tmp_path / "m.hcs")
    pm.save_hcs(hcs)
    outs = []
    for devs in ([0], list(range(torch.cuda.device_count()))):
        wd = tmp_path / ("run%d" % len(devs))
        wd.mkdir()
        cfg = Config({
            "restraints": {"Hi-C": {"input_matrix": hcs, "intra_sigma_list": [0.01], "inter_sigma_list": [0.01],
                                    "contact_range": 2.0, "tmp_dir": "actdist", "gpu_devices": devs}},
            "optimization": {"structure_output": hss, "iter_corr_knob": 1},
            "parameters": {"workdir": str(wd), "tmp_dir": str(wd / "tmp")},
            "runtime": {"Hi-C": {}}})
        ActivationDistanceStep(cfg).run()
        with hdf5.open_h5(cfg["runtime"]["Hi-C"]["actdist_file"]) as f:
            outs.append({k: f[k][()] for k in ("row", "col", "dist", "prob")})
    assert len(outs[0]["row"]) > 100
    for k in ("row", "col", "dist", "prob"):
        assert outs[0][k].tobytes() == outs[1][k].tobytes(), k


def test_dropin_runs_under_the_reference_step_run_on_the_gpu(tmp_path):
    """tests/real_step_driver.py with the real kernels: the drop-in driven by the reference's
    own Step.run / Config / SerialController / StepDB, records equal to the oracle's, second
    run skipped from the restart log."""
    import json
    import subprocess
    import sys
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("no copy of the reference package (oracle/_ref/igm)")
    r = subprocess.run([sys.executable, os.path.join(H.ROOT, "tests", "real_step_driver.py"), str(tmp_path)],
                       capture_output=True, text=True, timeout=600, cwd=H.ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["records"] > 0 and d["records_equal_oracle"] and not d["fake_gpu"]
    assert d["statuses"] == ["entry", "setup", "map", "mapped", "reduced", "cleanup", "completed"]
    assert d["task_calls_first_run"] == 1 and d["task_calls_second_run"] == 0
    assert d["second_run_actdist_file"] == d["actdist_file"]
