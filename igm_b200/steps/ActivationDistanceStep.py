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
device id), ``gpu_devices`` / ``gpu_max_devices`` (devices one task drives concurrently;
default: all visible), ``write_text_tmp`` (also write the reference's ``%d.out.tmp``),
``reference_task_files`` (``%d.in.npy`` in the reference's float64 (n, 4) layout instead of
one typed row per pair), ``write_task_files`` (default: only when the controller is not the
serial one - tasks of this process take their arrays from memory and leave their records there).
"""
from __future__ import annotations

import os
import shutil

import numpy as np

from .. import _lib, hdf5
from ..engine import ActdistEngine
from ..population import Population, ProbMatrix
from ._compat import Step, logger, make_absolute_path

# igm/steps/ActivationDistanceStep.py:32-38
actdist_shape = [("row", "int32"), ("col", "int32"), ("dist", "float32"), ("prob", "float32")]
actdist_fmt_str = "%6d %6d %10.4f %.4f"


def _pair_buffer(n, dtype, name):
    """Page-locked buffer of the A-step's pair list (tags shared with actdist_on_devices, which
    then hands the arrays to the device without another copy); plain memory without a GPU."""
    try:
        from ..engine import pinned_array
        return pinned_array((n,), dtype, ("actdist", name))
    except _lib.IgmkError:
        return np.empty(n, dtype)


def count_candidates(pm: ProbMatrix, intra_sigma, inter_sigma) -> int:
    """Length of the list ``filter_candidates`` would return (one counting pass in the C library)."""
    use_intra = intra_sigma is not False and intra_sigma is not None
    use_inter = inter_sigma is not False and inter_sigma is not None
    if not (use_intra or use_inter):
        return 0
    k = _lib.load().igmk_filter_candidates(pm.n, _lib.ptr(pm.indptr), _lib.ptr(pm.indices), _lib.ptr(pm.data),
                                           _lib.ptr(pm.chrom), int(use_intra),
                                           float(np.float32(intra_sigma)) if use_intra else 0.0, int(use_inter),
                                           float(np.float32(inter_sigma)) if use_inter else 0.0,
                                           None, None, None, 0)
    return int(-1 - k) if k < 0 else int(k)


def filter_candidates(pm: ProbMatrix, intra_sigma, inter_sigma, compare_dtype=np.float32, native=True,
                      pinned=False, capacity=None):
    """Candidate filter of the reference's setup loop (:166-178), vectorised.

    Stored non-zeros are visited in CSR row-major order (what ``coo_generator``
    yields).  ``pwish >= sigma`` is evaluated in ``compare_dtype``: the
    reference compares a float32 matrix value with a Python float, which under
    NumPy >= 2 happens in float32.  ``i == j`` entries are
    dropped (they never produce a record, :379-380).  Returns (i, j, pwish64).
    ``native=False`` (or a compare dtype other than float32) takes the NumPy form.
    ``pinned``: the native form writes into the step's page-locked pair buffers (sized by the
    matrix, so they are locked once per run) and returns views of them.
    """
    dt = np.dtype(compare_dtype)
    use_intra = intra_sigma is not False and intra_sigma is not None
    use_inter = inter_sigma is not False and inter_sigma is not None
    if not (use_intra or use_inter):
        z = np.zeros(0, np.int32)
        return z, z.copy(), np.zeros(0, np.float64)
    if dt == np.float32 and native:
        # one pass over the CSR arrays in the C library (igmk_filter_candidates)
        lib = _lib.load()
        cap = int(len(pm.indices)) if capacity is None else int(capacity)
        if pinned:
            oi, oj, op = _pair_buffer(cap, np.int32, "i"), _pair_buffer(cap, np.int32, "j"), _pair_buffer(cap, np.float64, "w")
        else:
            oi, oj, op = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.float64)
        k = lib.igmk_filter_candidates(pm.n, _lib.ptr(pm.indptr), _lib.ptr(pm.indices), _lib.ptr(pm.data),
                                       _lib.ptr(pm.chrom), int(use_intra),
                                       float(np.float32(intra_sigma)) if use_intra else 0.0, int(use_inter),
                                       float(np.float32(inter_sigma)) if use_inter else 0.0,
                                       _lib.ptr(oi), _lib.ptr(oj), _lib.ptr(op), cap)
        if k < 0 and capacity is not None:           # a capacity hint that was too small
            return filter_candidates(pm, intra_sigma, inter_sigma, compare_dtype, native, pinned, None)
        if k < 0:
            raise RuntimeError("igmk_filter_candidates failed (%d)" % k)
        if pinned:
            return oi[:k], oj[:k], op[:k]
        return oi[:k].copy(), oj[:k].copy(), op[:k].copy()
    pw = pm.data.astype(dt, copy=False)
    # first cut on the probability alone (the smaller threshold), then the intra / inter
    # distinction on the survivors only
    lo = min([dt.type(x) for x, u in ((intra_sigma, use_intra), (inter_sigma, use_inter)) if u])
    idx = np.flatnonzero(pw >= lo)
    rows = pm.rows()[idx]
    cols = pm.indices[idx].astype(np.int32, copy=False)
    pk = pw[idx]
    if not (use_intra and use_inter and dt.type(intra_sigma) == dt.type(inter_sigma)):
        intra = pm.chrom[rows] == pm.chrom[cols]
        keep = np.zeros(len(idx), dtype=bool)
        if use_intra:
            keep |= intra & (pk >= dt.type(intra_sigma))
        if use_inter:
            keep |= (~intra) & (pk >= dt.type(inter_sigma))
        keep &= rows != cols
    else:
        keep = rows != cols
    if not keep.all():
        idx, rows, cols = idx[keep], rows[keep], cols[keep]
    return (np.ascontiguousarray(rows), np.ascontiguousarray(cols),
            pm.data[idx].astype(np.float64))


def lookup_plast(last_actdist_file, n, ii, jj, native=True, out=None):
    """``plast[i, j]`` of setup (:144-160,177): the previous iteration's stored
    ``prob`` of the record whose (row, col) are the haploid indices themselves
    (only records with row < n and col < n survive the mask, quirk q7)."""
    if out is None:
        out = np.zeros(len(ii), dtype=np.float64)
    else:
        out = out[:len(ii)]
        out[:] = 0.0
    if last_actdist_file is None:
        return out
    with hdf5.open_h5(last_actdist_file) as h5f:
        row = np.asarray(h5f["row"][()])
        col = np.asarray(h5f["col"][()])
        prob = np.asarray(h5f["prob"][()])
    if native and len(ii) and n < 2 ** 31:
        # both sides in (row, col) order - the usual case: one merge pass in the C library
        r32, c32 = np.ascontiguousarray(row, np.int32), np.ascontiguousarray(col, np.int32)
        p32 = np.ascontiguousarray(prob, np.float32)
        i32, j32 = np.ascontiguousarray(ii, np.int32), np.ascontiguousarray(jj, np.int32)
        rc = _lib.load().igmk_join_plast(len(r32), _lib.ptr(r32), _lib.ptr(c32), _lib.ptr(p32), int(n), len(i32),
                                         _lib.ptr(i32), _lib.ptr(j32), _lib.ptr(out))
        if rc == 1:
            return out
        if rc < 0:
            raise RuntimeError("igmk_join_plast: bad arguments")
        out[:] = 0.0
    m = np.logical_and(row < n, col < n)
    key = row[m].astype(np.int64) * n + col[m].astype(np.int64)
    val = prob[m].astype(np.float32)
    if len(key) == 0:
        return out
    if len(key) > 1 and not np.all(key[1:] > key[:-1]):
        order = np.argsort(key, kind="stable")
        key, val = key[order], val[order]
        # duplicates (none in files this step writes) are summed, as coo -> lil does
        uk, start = np.unique(key, return_index=True)
        sums = np.add.reduceat(val, start).astype(np.float32)
    else:
        # files written by this step list the records in candidate (CSR) order: the
        # copy-0 x copy-0 keys are already strictly increasing
        uk, sums = key, val
    q = ii.astype(np.int64) * n + jj.astype(np.int64)
    if len(q) > 1 and len(uk) < len(q) and np.all(q[1:] > q[:-1]):
        # sorted candidate list (CSR order) and fewer stored records than candidates (the
        # usual case: the previous sigma was larger): look the records up in the list
        pos = np.searchsorted(q, uk)
        pos_c = np.minimum(pos, len(q) - 1)
        hit = q[pos_c] == uk
        out[pos_c[hit]] = sums[hit].astype(np.float64)
        return out
    pos = np.searchsorted(uk, q)
    pos_c = np.minimum(pos, len(uk) - 1)
    hit = uk[pos_c] == q
    out[hit] = sums[pos_c[hit]].astype(np.float64)
    return out


def pack_records(row, col, dist, prob):
    """Task output: the four record columns as the rows of one (4, n) 32-bit array
    (columnar, so reduce() hands them to the HDF5 writer without re-interleaving)."""
    out = np.empty((4, len(row)), dtype=np.uint32)
    out[0] = np.asarray(row, np.int32).view(np.uint32)
    out[1] = np.asarray(col, np.int32).view(np.uint32)
    out[2] = np.asarray(dist, np.float32).view(np.uint32)
    out[3] = np.asarray(prob, np.float32).view(np.uint32)
    return out


def unpack_records(parts):
    """Columns {row, col: int32; dist, prob: float32} of the concatenated task outputs
    (:246-257).  Accepts the columnar arrays of pack_records and structured
    ``actdist_shape`` arrays (what np.genfromtxt yields in the reference)."""
    cols = []
    for a in parts:
        if a.dtype.names:
            cols.append(pack_records(a["row"], a["col"], a["dist"], a["prob"]))
        elif a.size == 0:
            cols.append(np.zeros((4, 0), np.uint32))
        else:
            cols.append(a)
    rec = cols[0] if len(cols) == 1 else (np.concatenate(cols, axis=1) if cols else np.zeros((4, 0), np.uint32))
    return {"row": rec[0].view(np.int32), "col": rec[1].view(np.int32),
            "dist": rec[2].view(np.float32), "prob": rec[3].view(np.float32)}


_engine_cache = {}
_staged_cache = {}
_matrix_cache = {}
_handoff = {}          # in-file path -> (ii, jj, pw, pl): setup() -> task() inside one process
_handoff_out = {}      # out-file path -> packed records: task() -> reduce() inside one process
_max_pairs = {}        # tmp_dir -> stored entries of the matrix = the longest list any sigma can produce
LAST_TIMING = {}       # seconds spent in the phases of the most recent setup / task / reduce (diagnostic)

# task input file: one row per candidate pair (the reference stores a float64 (n, 4) array)
PAIR_DTYPE = np.dtype([("i", np.int32), ("j", np.int32), ("pwish", np.float64), ("plast", np.float64)])


def _file_key(path):
    st = os.stat(path)
    return (os.path.abspath(path), st.st_mtime_ns, st.st_size)


def _load_matrix(path):
    """Parsed .hcs, kept across the sigma iterations of a run (the matrix does not change;
    the reference re-reads it in every setup, :124-126)."""
    key = _file_key(path)
    pm = _matrix_cache.get(key)
    if pm is None:
        _matrix_cache.clear()
        pm = ProbMatrix.from_hcs(path)
        _matrix_cache[key] = pm
    return pm


def _stage_population(hss_path):
    from ..engine import StagedHss
    return StagedHss(hss_path)


def _get_engines(hss_path, devices, defer_upload=False):
    """One engine per device for the population file (keyed by path, mtime, size): the
    coordinates are read ONCE into pinned host memory and staged into every device's HBM
    concurrently; engines of an older population are closed.  ``defer_upload``: new engines
    keep the coordinates in host memory for ``ActdistEngine.actdist_buffers`` to stage inside
    its pipelined device call (every A-step follows an M-step that rewrote the file, so a new
    population per step is the normal case)."""
    from concurrent.futures import ThreadPoolExecutor
    key = _file_key(hss_path)
    for k in list(_engine_cache):
        if k[:3] != key:
            _engine_cache.pop(k).close()
    for k in list(_staged_cache):
        if k != key:
            _staged_cache.pop(k)
    missing = [d for d in devices if key + (d,) not in _engine_cache]
    if missing:
        st = _staged_cache.get(key)
        if st is None:
            st = _staged_cache[key] = _stage_population(hss_path)
        if len(missing) == 1:
            kw = {"upload": False} if defer_upload else {}
            _engine_cache[key + (missing[0],)] = ActdistEngine.from_hss(hss_path, missing[0], staged=st, **kw)
        else:
            # several devices: each uploads only its share of the beads from host memory, then
            # pulls the other shares from its peers over NVLink - one population's worth of
            # PCIe traffic in total instead of one per device
            from ..dist import bead_shares
            _, shares = bead_shares(st.nbead, len(missing))
            with ThreadPoolExecutor(len(missing)) as ex:
                new = list(ex.map(lambda d: ActdistEngine.from_hss(hss_path, d, staged=st, upload=False), missing))

                def own(k):
                    lo, hi = shares[k]
                    new[k]._pending_xyz = None
                    if hi > lo:
                        new[k].upload_coordinates(st.coordinates[lo:hi], bead0=lo)

                def pull(k):
                    for step in range(1, len(new)):              # staggered: no two devices read one peer at once
                        src = (k + step) % len(new)
                        lo, hi = shares[src]
                        if hi > lo:
                            new[k].copy_coordinates_from(new[src], lo, hi - lo)
                list(ex.map(own, range(len(new))))
                list(ex.map(pull, range(len(new))))
            for d, eng in zip(missing, new):
                _engine_cache[key + (d,)] = eng
    engines = [_engine_cache[key + (d,)] for d in devices]
    if not defer_upload:
        for eng in engines:
            if getattr(eng, "_pending_xyz", None) is not None:
                eng.stage_pending()
    return engines


def _get_engine(hss_path, device):
    return _get_engines(hss_path, [device])[0]


def visible_devices(dictHiC):
    """Devices one task may drive: ``gpu_devices`` (list) or every visible GPU from
    ``gpu_device`` on, at most ``gpu_max_devices``."""
    if dictHiC.get("gpu_devices"):
        return [int(d) for d in dictHiC["gpu_devices"]]
    import torch  # device enumeration only
    ndev = torch.cuda.device_count() if torch.cuda.is_available() else 0
    first = int(dictHiC.get("gpu_device", 0))
    devs = list(range(first, max(ndev, first + 1)))
    return devs[:int(dictHiC.get("gpu_max_devices", len(devs)))] or [first]


def actdist_on_devices(hss_path, devices, ii, jj, pw, pl, contact_range, it_corr, mode, max_pairs=0):
    """get_actdist for every pair, the list cut into contiguous shares, one per device, all
    devices working at the same time (one host thread per device; the C library releases
    the GIL).  The reference runs its batches concurrently on the workers of its controller
    (igm/core/step.py:259-274, igm/parallel/ipyparallel_controller.py:66-109); with the
    serial controller this is what gives igm-run all the GPUs of the box."""
    from concurrent.futures import ThreadPoolExecutor
    from ..engine import is_pinned_view, pinned_array
    import time
    n = len(ii)
    t0 = time.perf_counter()
    engines = _get_engines(hss_path, devices, defer_upload=True)
    LAST_TIMING["task_engines_s"] = time.perf_counter() - t0
    bounds = np.linspace(0, n, len(devices) + 1).astype(np.int64)
    t0 = time.perf_counter()
    # ONE page-locked copy of the list and of the results for all devices (page-locking is
    # serialised by the kernel: per-device buffers made an 8-GPU task slower than a 1-GPU one);
    # every device works on its contiguous share of them
    # (setup() in this process already wrote the list there: nothing to copy then)
    def staged(arr, dtype, name):
        if arr.dtype == dtype and is_pinned_view(arr, ("actdist", name)):
            return arr
        buf = pinned_array((n,), dtype, ("actdist", name))
        buf[:] = arr
        return buf
    p_i, p_j = staged(ii, np.int32, "i"), staged(jj, np.int32, "j")
    p_w, p_l = staged(pw, np.float64, "w"), staged(pl, np.float64, "l")
    out = pinned_array((n,), _lib.PAIR_RESULT_DTYPE, ("actdist", "o"))
    LAST_TIMING["task_pinned_buffers_s"] = time.perf_counter() - t0
    t1 = time.perf_counter()

    def run(k):
        lo, hi = int(bounds[k]), int(bounds[k + 1])
        if max_pairs:                       # the longest list of the run: device buffers sized once
            engines[k].reserve_pairs(-(-int(max_pairs) // len(devices)))
        if hi == lo:
            engines[k].stage_pending()
            return
        engines[k].actdist_buffers(hi - lo, p_i[lo:hi], p_j[lo:hi], p_w[lo:hi], p_l[lo:hi], out[lo:hi],
                                   contact_range, it_corr, str(mode).upper())
    if len(devices) == 1:
        run(0)
    else:
        with ThreadPoolExecutor(len(devices)) as ex:
            list(ex.map(run, range(len(devices))))
    LAST_TIMING["task_library_call_s"] = time.perf_counter() - t1
    LAST_TIMING["task_device_s"] = time.perf_counter() - t0
    return engines[0], out


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

        pm = _load_matrix(dictHiC["input_matrix"])
        n = pm.n
        last_actdist_file = self.cfg.get("runtime/Hi-C").get("actdist_file", None)
        n_shards = max(1, int(dictHiC.get("gpu_shards", 1)))

        self.tmp_extensions = [".npy", ".tmp"]
        self.tmp_dir = make_absolute_path(
            self.cfg.get("restraints/Hi-C/tmp_dir", "actdist"),
            self.cfg.get("parameters/tmp_dir"))
        self.keep_temporary_files = dictHiC.get("keep_temporary_files", False)
        os.makedirs(self.tmp_dir, exist_ok=True)

        import time
        t0 = time.perf_counter()
        # the longest list this run can still produce (the smallest sigmas of the remaining
        # schedule, :69-98): the page-locked buffers and the device buffers are sized for it once
        def smallest(cur, rest):
            vals = [v for v in [cur] + list(rest or []) if v is not False and v is not None]
            return min(vals) if vals else cur
        cap = count_candidates(pm, smallest(intra_sigma, self.cfg.get("runtime/Hi-C/intra_sigma_list", [])),
                               smallest(inter_sigma, self.cfg.get("runtime/Hi-C/inter_sigma_list", [])))
        ii, jj, pw = filter_candidates(pm, intra_sigma, inter_sigma, pinned=True, capacity=cap)
        LAST_TIMING["setup_filter_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        pl = lookup_plast(last_actdist_file, n, ii, jj, out=_pair_buffer(cap, np.float64, "l"))
        _pair_buffer(cap, _lib.PAIR_RESULT_DTYPE, "o")     # the task's result buffer: locked once per run too
        LAST_TIMING["setup_plast_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        # contiguous, equal-count shards keep bead-i locality and output order.  The task input
        # is one typed row per pair (the reference writes float64 (n, 4)); a task that runs in
        # this process takes the arrays from memory instead of reading the file back
        bounds = np.linspace(0, len(ii), n_shards + 1).astype(np.int64)
        _handoff.clear()
        _handoff_out.clear()
        # a serial controller runs the tasks inside this process (igm/parallel/parallel_controller.py):
        # the arrays are handed over in memory and the task files are not needed
        in_process = "Serial" in type(getattr(self, "controller", None)).__name__
        write_files = (bool(dictHiC.get("write_task_files", not in_process)) or bool(self.keep_temporary_files)
                       or bool(dictHiC.get("reference_task_files", False)) or bool(dictHiC.get("write_text_tmp", False)))
        for b in range(n_shards):
            lo, hi = bounds[b], bounds[b + 1]
            fname = os.path.join(self.tmp_dir, "%d.in.npy" % b)
            if dictHiC.get("reference_task_files", False):
                # the reference's own layout: float64 (n, 4) = (i, j, pwish, plast) (:181-186)
                np.save(fname, np.stack([ii[lo:hi].astype(np.float64), jj[lo:hi].astype(np.float64),
                                         pw[lo:hi], pl[lo:hi]], axis=1))
            elif write_files:
                rows = np.empty(hi - lo, dtype=PAIR_DTYPE)
                rows["i"], rows["j"], rows["pwish"], rows["plast"] = ii[lo:hi], jj[lo:hi], pw[lo:hi], pl[lo:hi]
                np.save(fname, rows)
            _handoff[fname] = (ii[lo:hi], jj[lo:hi], pw[lo:hi], pl[lo:hi], write_files)
        LAST_TIMING["setup_files_s"] = time.perf_counter() - t0
        self.argument_list = range(n_shards)
        self.n_candidate_pairs = int(len(ii))
        _max_pairs.clear()
        _max_pairs[self.tmp_dir] = cap // n_shards + 1

    @staticmethod
    def task(batch_id, cfg, tmp_dir):
        dictHiC = cfg["restraints"]["Hi-C"]
        it_corr = cfg.get("runtime/Hi-C/iter_corr_knob")
        in_name = os.path.join(tmp_dir, "%d.in.npy" % batch_id)
        out_name = os.path.join(tmp_dir, "%d.out.npy" % batch_id)
        held = _handoff.pop(in_name, None)
        write_out = True
        if held is not None:
            ii, jj, pw, pl, write_out = held
        else:
            params = np.load(in_name, mmap_mode="r")
            if params.dtype.names:
                ii, jj, pw, pl = params["i"], params["j"], params["pwish"], params["plast"]
            elif params.size:                                # the reference's float64 (n, 4) layout
                ii, jj, pw, pl = (params[:, 0].astype(np.int32), params[:, 1].astype(np.int32),
                                  params[:, 2], params[:, 3])
            else:
                ii = np.zeros(0, np.int32)
                jj, pw, pl = ii, np.zeros(0), np.zeros(0)
        if len(ii) == 0:
            if write_out:
                np.save(out_name, np.zeros((4, 0), dtype=np.uint32))
            else:
                _handoff_out[out_name] = np.zeros((4, 0), dtype=np.uint32)
            return
        # one task drives every visible GPU at once; several tasks (gpu_shards > 1: workers of
        # a parallel controller) take one device each
        devices = visible_devices(dictHiC)
        n_shards = max(1, int(dictHiC.get("gpu_shards", 1)))
        if n_shards > 1:
            devices = [devices[batch_id % len(devices)]]
        eng, res = actdist_on_devices(cfg.get("optimization/structure_output"), devices, ii, jj, pw, pl,
                                      dictHiC.get("contact_range", 2.0), 1 if it_corr == 1 else 0,
                                      dictHiC.get("gpu_mode", "LB"), max_pairs=_max_pairs.get(tmp_dir, 0))
        import time
        t0 = time.perf_counter()
        row, col, dist, prob = eng.expand_records(ii, jj, res)
        LAST_TIMING["task_expand_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        if write_out:
            np.save(out_name, pack_records(row, col, dist, prob))
        else:
            _handoff_out[out_name] = (row, col, dist, prob)          # reduce() of this process takes them
        LAST_TIMING["task_save_s"] = time.perf_counter() - t0
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
        parts = []
        for i in self.argument_list:
            name = os.path.join(self.tmp_dir, "%d.out.npy" % i)
            held = _handoff_out.pop(name, None)                      # a task of this process left them in memory
            parts.append(held if held is not None else np.load(name))
        if len(parts) == 1 and isinstance(parts[0], tuple):
            columns = dict(zip(("row", "col", "dist", "prob"), parts[0]))
        else:
            columns = unpack_records([pack_records(*a) if isinstance(a, tuple) else a for a in parts])

        additional_data = []
        if "Hi-C" in self.cfg["runtime"]:
            additional_data.append("INTERsigma_{:.4f}".format(self.cfg["runtime"]["Hi-C"].get("inter_sigma", -1.0)))
            additional_data.append("INTRAsigma_{:.4f}".format(self.cfg["runtime"]["Hi-C"].get("intra_sigma", -1.0)))
        if "opt_iter" in self.cfg["runtime"]:
            additional_data.append("iter_{}".format(self.cfg["runtime"]["opt_iter"] - 1))

        tmp_actdist_file = actdist_file + ".tmp"
        hdf5.write_h5(tmp_actdist_file, columns)

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
