"""Drop-in replacement of IGM's Hi-C ``ActivationDistanceStep`` on B200.

Mirrors the operator interface of the reference step
(igm/steps/ActivationDistanceStep.py:41-309): same class name, ``name()``
format (it feeds the restart uid, igm/core/step.py:184), ``setup`` / static
``task(batch_id, cfg, tmp_dir)`` / ``reduce`` / ``skip``, the same config keys
(``restraints/Hi-C/{input_matrix, intra_sigma_list, inter_sigma_list,
contact_range, tmp_dir, keep_temporary_files, batch_size}``,
``optimization/{structure_output, iter_corr_knob}``, ``runtime/Hi-C/*``), the
same temp-file names, the same ``actdist.hdf5`` (``row,col: int32``,
``dist,prob: float32``) and the same swap-file naming.  ``igm-run`` calls it as
``igm.ActivationDistanceStep(cfg).run()`` (bin/igm-run:138-140,164-166).

What differs is where the work happens: ``task`` hands a whole shard of
candidate pairs to the CUDA library (igm_b200/csrc, through ctypes) instead of
looping ``get_actdist`` over 1000-pair batches, and values travel as binary
arrays - the 4-decimal text round trip of the reference (:38,:230,:249) is
reproduced numerically on the device, bit for bit.

Extra, optional keys (all under ``restraints/Hi-C``): ``gpu_mode`` ("LB" |
"GP"), ``gpu_shards`` (number of tasks, default 1), ``gpu_device`` (first
device id), ``write_text_tmp`` (also write the reference's ``%d.out.tmp``).
"""
from __future__ import annotations

import os
import shutil

import numpy as np

from .. import hdf5
from ..engine import ActdistEngine
from ..population import Population, ProbMatrix
from ._compat import Step, logger, make_absolute_path

# igm/steps/ActivationDistanceStep.py:32-38
actdist_shape = [("row", "int32"), ("col", "int32"), ("dist", "float32"), ("prob", "float32")]
actdist_fmt_str = "%6d %6d %10.4f %.4f"


def filter_candidates(pm: ProbMatrix, intra_sigma, inter_sigma, compare_dtype=np.float32):
    """Candidate filter of the reference's setup loop (:166-178), vectorised.

    Stored non-zeros are visited in CSR row-major order (what ``coo_generator``
    yields).  ``pwish >= sigma`` is evaluated in ``compare_dtype``: the
    reference compares a float32 matrix value with a Python float, which under
    NumPy >= 2 happens in float32.  ``i == j`` entries are
    dropped (they never produce a record, :379-380).  Returns (i, j, pwish64).
    """
    rows = pm.rows().astype(np.int64)
    cols = pm.indices.astype(np.int64)
    dt = np.dtype(compare_dtype)
    pw = pm.data.astype(dt)
    intra = pm.chrom[rows] == pm.chrom[cols]
    keep = np.zeros(len(rows), dtype=bool)
    if intra_sigma is not False and intra_sigma is not None:
        keep |= intra & (pw >= dt.type(intra_sigma))
    if inter_sigma is not False and inter_sigma is not None:
        keep |= (~intra) & (pw >= dt.type(inter_sigma))
    keep &= rows != cols
    return (rows[keep].astype(np.int32), cols[keep].astype(np.int32),
            pm.data[keep].astype(np.float64))


def lookup_plast(last_actdist_file, n, ii, jj):
    """``plast[i, j]`` of setup (:144-160,177): the previous iteration's stored
    ``prob`` of the record whose (row, col) are the haploid indices themselves
    (only records with row < n and col < n survive the mask, quirk q7)."""
    out = np.zeros(len(ii), dtype=np.float64)
    if last_actdist_file is None:
        return out
    with hdf5.open_h5(last_actdist_file) as h5f:
        row = np.asarray(h5f["row"][()])
        col = np.asarray(h5f["col"][()])
        prob = np.asarray(h5f["prob"][()])
    m = np.logical_and(row < n, col < n)
    key = row[m].astype(np.int64) * n + col[m].astype(np.int64)
    val = prob[m].astype(np.float32)
    if len(key) == 0:
        return out
    order = np.argsort(key, kind="stable")
    key, val = key[order], val[order]
    # duplicates (none in files this step writes) are summed, as coo -> lil does
    uk, start = np.unique(key, return_index=True)
    sums = np.add.reduceat(val, start).astype(np.float32)
    q = ii.astype(np.int64) * n + jj.astype(np.int64)
    pos = np.searchsorted(uk, q)
    pos_c = np.minimum(pos, len(uk) - 1)
    hit = uk[pos_c] == q
    out[hit] = sums[pos_c[hit]].astype(np.float64)
    return out


_engine_cache = {}


def _get_engine(hss_path, device):
    """One engine per (file, mtime, device) per process: coordinates are staged
    into HBM once per A-step, not once per batch."""
    st = os.stat(hss_path)
    key = (os.path.abspath(hss_path), st.st_mtime_ns, st.st_size, device)
    eng = _engine_cache.get(key)
    if eng is None:
        for k in list(_engine_cache):
            _engine_cache.pop(k).close()
        eng = ActdistEngine.from_hss(hss_path, device)     # chunk-wise staging, no host copy
        _engine_cache[key] = eng
    return eng


class ActivationDistanceStep(Step):

    def __init__(self, cfg):
        # identical runtime bookkeeping to the reference (:42-100)
        if "intra_sigma_list" not in cfg["runtime"]["Hi-C"]:
            cfg["runtime"]["Hi-C"]["intra_sigma_list"] = cfg["restraints"]["Hi-C"]["intra_sigma_list"][:]
        if "inter_sigma_list" not in cfg["runtime"]["Hi-C"]:
            cfg["runtime"]["Hi-C"]["inter_sigma_list"] = cfg["restraints"]["Hi-C"]["inter_sigma_list"][:]
        if "iter_corr_knob" not in cfg.get("runtime/Hi-C"):
            # quirk q1: the key is absent from the schema and the demo config
            cfg["runtime"]["Hi-C"]["iter_corr_knob"] = cfg.get("optimization/iter_corr_knob", 0)
        if ("inter_sigma" not in cfg["runtime"]["Hi-C"]) and ("intra_sigma" not in cfg["runtime"]["Hi-C"]):
            inters = cfg.get("runtime/Hi-C/inter_sigma_list")
            intras = cfg.get("runtime/Hi-C/intra_sigma_list")
            if len(inters) and len(intras):
                cfg.set("runtime/Hi-C/inter_sigma", inters.pop(0))
                cfg.set("runtime/Hi-C/intra_sigma", intras.pop(0))
        super(ActivationDistanceStep, self).__init__(cfg)

    def name(self):
        s = "ActivationDistanceStep (INTER sigma={:.2f}%, INTRA sigma={:.2f}%, iter={:s})"
        return s.format(
            self.cfg.get("runtime/Hi-C/inter_sigma") * 100.0,
            self.cfg.get("runtime/Hi-C/intra_sigma") * 100.0,
            str(self.cfg.get("runtime/opt_iter", "NA")))

    def setup(self):
        dictHiC = self.cfg["restraints"]["Hi-C"]
        inter_sigma = self.cfg.get("runtime/Hi-C/inter_sigma", False)
        intra_sigma = self.cfg.get("runtime/Hi-C/intra_sigma", False)
        logger.info(inter_sigma)
        logger.info(intra_sigma)

        pm = ProbMatrix.from_hcs(dictHiC["input_matrix"])
        n = pm.n
        last_actdist_file = self.cfg.get("runtime/Hi-C").get("actdist_file", None)
        n_shards = max(1, int(dictHiC.get("gpu_shards", 1)))

        self.tmp_extensions = [".npy", ".tmp"]
        self.tmp_dir = make_absolute_path(
            self.cfg.get("restraints/Hi-C/tmp_dir", "actdist"),
            self.cfg.get("parameters/tmp_dir"))
        self.keep_temporary_files = dictHiC.get("keep_temporary_files", False)
        os.makedirs(self.tmp_dir, exist_ok=True)

        ii, jj, pw = filter_candidates(pm, intra_sigma, inter_sigma)
        pl = lookup_plast(last_actdist_file, n, ii, jj)
        params = np.stack([ii.astype(np.float64), jj.astype(np.float64), pw, pl], axis=1)
        # contiguous, equal-count shards keep bead-i locality and output order
        bounds = np.linspace(0, len(ii), n_shards + 1).astype(np.int64)
        for b in range(n_shards):
            np.save(os.path.join(self.tmp_dir, "%d.in.npy" % b), params[bounds[b]:bounds[b + 1]])
        self.argument_list = range(n_shards)

    @staticmethod
    def task(batch_id, cfg, tmp_dir):
        dictHiC = cfg["restraints"]["Hi-C"]
        it_corr = cfg.get("runtime/Hi-C/iter_corr_knob")
        params = np.load(os.path.join(tmp_dir, "%d.in.npy" % batch_id))
        out_name = os.path.join(tmp_dir, "%d.out.npy" % batch_id)
        if params.size == 0:
            np.save(out_name, np.zeros(0, dtype=actdist_shape))
            return
        import torch  # device enumeration only
        ndev = torch.cuda.device_count() if torch.cuda.is_available() else 0
        device = int(dictHiC.get("gpu_device", 0)) + (batch_id % max(1, ndev))
        eng = _get_engine(cfg.get("optimization/structure_output"), device)
        ii = params[:, 0].astype(np.int32)
        jj = params[:, 1].astype(np.int32)
        res = eng.actdist(ii, jj, params[:, 2], params[:, 3],
                          contact_range=dictHiC.get("contact_range", 2.0),
                          it_corr=1 if it_corr == 1 else 0,
                          mode=dictHiC.get("gpu_mode", "LB"))
        row, col, dist, prob = eng.expand_records(ii, jj, res)
        rec = np.empty(len(row), dtype=actdist_shape)
        rec["row"], rec["col"], rec["dist"], rec["prob"] = row, col, dist, prob
        np.save(out_name, rec)
        if dictHiC.get("write_text_tmp", False):
            # the reference's own wire format (:228-230), for byte-level comparison
            ad = np.repeat(np.sqrt(res["d2_sel_bits"].view(np.float32).astype(np.float64)), res["nrec"])
            pp = np.repeat(res["p"], res["nrec"])
            with open(os.path.join(tmp_dir, "%d.out.tmp" % batch_id), "w") as f:
                f.write("\n".join([actdist_fmt_str % x for x in
                                   zip(row.tolist(), col.tolist(), ad.tolist(), pp.tolist())]))

    def reduce(self):
        actdist_file = os.path.join(self.tmp_dir, "actdist.hdf5")
        last_actdist_file = self.cfg["runtime"]["Hi-C"].get("actdist_file", None)
        parts = [np.load(os.path.join(self.tmp_dir, "%d.out.npy" % i)) for i in self.argument_list]
        rec = np.concatenate(parts) if parts else np.zeros(0, dtype=actdist_shape)

        additional_data = []
        if "Hi-C" in self.cfg["runtime"]:
            additional_data.append("INTERsigma_{:.4f}".format(self.cfg["runtime"]["Hi-C"].get("inter_sigma", -1.0)))
            additional_data.append("INTRAsigma_{:.4f}".format(self.cfg["runtime"]["Hi-C"].get("intra_sigma", -1.0)))
        if "opt_iter" in self.cfg["runtime"]:
            additional_data.append("iter_{}".format(self.cfg["runtime"]["opt_iter"] - 1))

        tmp_actdist_file = actdist_file + ".tmp"
        hdf5.write_h5(tmp_actdist_file, {
            "row": np.ascontiguousarray(rec["row"]), "col": np.ascontiguousarray(rec["col"]),
            "dist": np.ascontiguousarray(rec["dist"]), "prob": np.ascontiguousarray(rec["prob"])})

        swapfile = os.path.realpath(".".join([actdist_file, ] + additional_data))
        if last_actdist_file is not None:
            shutil.move(last_actdist_file, swapfile)
        shutil.move(tmp_actdist_file, actdist_file)
        self.cfg["runtime"]["Hi-C"]["actdist_file"] = actdist_file

    def skip(self):
        self.tmp_dir = make_absolute_path(
            self.cfg.get("restraints/Hi-C/tmp_dir", "actdist"),
            self.cfg.get("parameters/tmp_dir"))
        self.actdist_file = os.path.join(self.tmp_dir, "actdist.hdf5")
        self.cfg["runtime"]["Hi-C"]["actdist_file"] = self.actdist_file
