#!/usr/bin/env python
"""TEST DRIVER - runs the Step drop-in under the reference's REAL ``igm.core.Step.run``
(igm/core/step.py:226-322: sqlite restart log of igm/core/job_tracking.py, skip() on a
re-run) with the reference's own ``Config`` and ``SerialController``.

    python tests/real_step_driver.py WORKDIR [--fake-gpu]

The reference package is loaded through oracle/ref_loader.py (stubs for the third-party
imports it does not need here) BEFORE igm_b200.steps is imported, so that
igm_b200/steps/_compat.py binds the drop-in to the reference's Step class - exactly what
happens in a production IGM install.  --fake-gpu (CPU test box): the device call is
replaced by the NumPy oracle; everything else (setup, files, reduce, restart log) is real.
Prints one JSON line.
"""
import json
import os
import sqlite3
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    work = sys.argv[1]
    fake = "--fake-gpu" in sys.argv
    os.environ.setdefault("HOME", work)
    from oracle import ref_loader
    from oracle import actdist_oracle as orc
    ref_loader.install()
    import igm                                             # noqa: F401  (the reference package)
    from igm.core import Config as RefConfig, Step as RefStep
    from igm.parallel import SerialController
    import igm_b200.steps._compat as compat
    import igm_b200.steps                                  # noqa: F401
    S = sys.modules["igm_b200.steps.ActivationDistanceStep"]   # (the package re-exports the class under that name)
    from igm_b200 import hdf5, synthetic
    assert compat.HAVE_REFERENCE_STEP and compat.Step is RefStep
    assert issubclass(S.ActivationDistanceStep, RefStep)

    pop = synthetic.make_population(2_000_000, 60, seed=4, genome_scale=0.02)
    hss = os.path.join(work, "pop.hss")
    pop.save_hss(hss)
    pm = synthetic.make_prob_matrix(pop.chrom_hap(), seed=9, inter_per_row=6.0)
    hcs = os.path.join(work, "m.hcs")
    pm.save_hcs(hcs)
    calls = {"task": 0}
    if fake:
        def fake_devices(hss_path, devices, ii, jj, pw, pl, contact_range, it_corr, mode, max_pairs=0):
            calls["task"] += 1
            recs, dets = orc.run_pairs(ii, jj, pw, pl, pop.coordinates, pop.radii, pop.chrom_hap(), pop.copy_index,
                                       it_corr, contact_range, orc.MODE_LB)
            cols = orc.records_to_arrays(recs)

            class Eng:
                def expand_records(self, i, j, res):
                    return cols
            return Eng(), None
        S.actdist_on_devices = fake_devices
        S.visible_devices = lambda d: [0]
    else:
        real = S.actdist_on_devices

        def counted(*a, **k):
            calls["task"] += 1
            return real(*a, **k)
        S.actdist_on_devices = counted

    def make_cfg():
        return RefConfig({
            "parameters": {"workdir": work, "tmp_dir": "tmp", "step_db": os.path.join(work, "stepdb.sqlite")},
            "parallel": {"controller": "serial"},
            "optimization": {"structure_output": hss, "iter_corr_knob": 1, "optimizer_options": {}},
            "restraints": {"Hi-C": {"input_matrix": hcs, "intra_sigma_list": [0.05], "inter_sigma_list": [0.05],
                                    "contact_range": 2.0, "tmp_dir": "actdist"}}})
    cfg = make_cfg()
    step = S.ActivationDistanceStep(cfg)
    assert isinstance(step.controller, SerialController)
    step.run()                                             # the reference's Step.run, not ours
    out = cfg["runtime"]["Hi-C"]["actdist_file"]
    with hdf5.open_h5(out) as f:
        got = {k: f[k][()] for k in ("row", "col", "dist", "prob")}
    ii, jj, pw = S.filter_candidates(pm, 0.05, 0.05, native=not fake)
    recs, _ = orc.run_pairs(ii, jj, pw, np.zeros(len(ii)), pop.coordinates, pop.radii, pop.chrom_hap(),
                            pop.copy_index, 1, 2.0, orc.MODE_LB)
    row, col, dist, prob = orc.records_to_arrays(recs)
    same = bool(np.array_equal(got["row"], row) and np.array_equal(got["col"], col)
                and np.array_equal(got["dist"].view(np.uint32), dist.view(np.uint32))
                and np.array_equal(got["prob"].view(np.uint32), prob.view(np.uint32)))
    with sqlite3.connect(os.path.join(work, "stepdb.sqlite")) as conn:
        statuses = [r[0] for r in conn.execute("SELECT status FROM steps ORDER BY time, rowid").fetchall()]
    first_calls = calls["task"]

    # second run of the same step from a fresh configuration: the restart log says
    # "completed" -> Step.run restores the runtime and calls skip() (igm/core/step.py:245-252)
    cfg2 = make_cfg()
    step2 = S.ActivationDistanceStep(cfg2)
    assert step2.uid == step.uid
    step2.run()
    print(json.dumps({
        "records": int(len(row)), "records_equal_oracle": same, "statuses": statuses,
        "task_calls_first_run": first_calls, "task_calls_second_run": calls["task"] - first_calls,
        "second_run_actdist_file": cfg2["runtime"]["Hi-C"].get("actdist_file"), "actdist_file": out,
        "second_run_sigma": cfg2["runtime"]["Hi-C"].get("inter_sigma"),
        "fake_gpu": fake, "reference_root": ref_loader.REFERENCE_ROOT}))


if __name__ == "__main__":
    main()
