"""Host-side driver of the CUDA library for one population on one GPU.

``ActdistEngine`` is what the Step drop-in (igm_b200/steps) and bench.py call;
it owns one ``igmk_ctx`` (coordinates resident in HBM, bead-major,
structure-contiguous) and exposes the hot path of the reference's
``ActivationDistanceStep.task`` loop (igm/steps/ActivationDistanceStep.py:
215-222) plus the contact-frequency tiles.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import (ALGO_FAST, ALGO_SIMPLE, MODE_GP, MODE_LB, PAIR_RESULT_DTYPE, IgmkError,
                   check, ptr)
from .population import Population

import threading

_pinned_cache = {}     # key -> (address, ndarray view): page-locked staging buffers, kept per process
_pinned_lock = threading.Lock()   # one task drives several devices from several threads


def pinned_array(shape, dtype=np.float32, tag=None) -> np.ndarray:
    """Page-locked host array (cudaHostAlloc through igmk_host_alloc) - asynchronous H2D /
    D2H copies need it.  Buffers are kept for the life of the process, one per ``tag`` (a
    tagged buffer only ever grows, by at least 1.5x, and serves every smaller request: the
    candidate lists of a sigma sweep grow from step to step) or one per size (no tag):
    page-locking runs at 1 - 3 GB/s, far too slow to repeat for every A-step."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    key = ("size", nbytes) if tag is None else ("tag", tag)
    with _pinned_lock:
        ent = _pinned_cache.get(key)
        if ent is None or len(ent[1]) < nbytes:
            lib = _lib.load()
            cap = max(nbytes, 1)
            if ent is not None:
                cap = max(cap, int(1.5 * len(ent[1])))
                lib.igmk_host_free(C.c_void_p(ent[0]))
                del _pinned_cache[key]
            p = C.c_void_p()
            check(lib.igmk_host_alloc(C.byref(p), cap))
            buf = (C.c_char * cap).from_address(p.value)
            ent = (p.value, np.frombuffer(buf, dtype=np.uint8, count=cap))
            # bounded: drop the oldest size-keyed buffer (a population of an earlier step);
            # tagged buffers belong to a device's working set and stay
            sized = [k for k in _pinned_cache if k[0] == "size"]
            if len(sized) >= 8:
                lib.igmk_host_free(C.c_void_p(_pinned_cache.pop(sized[0])[0]))
            _pinned_cache[key] = ent
    return ent[1][:nbytes].view(dtype).reshape(shape)


def is_pinned_view(arr, tag) -> bool:
    """True when ``arr`` is a view that starts at the base of the tagged page-locked buffer
    (so a consumer can hand it to the device as it is instead of copying it there)."""
    ent = _pinned_cache.get(("tag", tag))
    return (ent is not None and isinstance(arr, np.ndarray) and arr.flags["C_CONTIGUOUS"]
            and arr.ctypes.data == ent[0] and arr.nbytes <= len(ent[1]))


class StagedHss:
    """Host-side staging of a population file: coordinates in pinned memory + the index
    tables ``get_actdist`` needs (igm/steps/ActivationDistanceStep.py:382-393)."""

    def __init__(self, path: str, threads: int = 8):
        import json
        import os
        from concurrent.futures import ThreadPoolExecutor
        from . import hdf5
        from .population import CopyIndex
        with hdf5.open_h5(path) as f:
            ds = f["coordinates"]
            self.nbead, self.nstruct = int(ds.shape[0]), int(ds.shape[1])
            ext = ds.raw_extents() if hasattr(ds, "raw_extents") else None
            if ext is not None and str(ds.dtype) in ("float32", "<f4"):
                crd = pinned_array((self.nbead, self.nstruct, 3), np.float32)
                flat = crd.reshape(-1).view(np.uint8)
                piece = 16 << 20
                jobs = [(fo + a, min(piece, nb - a), do + a) for fo, nb, do in ext for a in range(0, nb, piece)]
                fd = os.open(path, os.O_RDONLY)
                try:
                    def rd(job):
                        fo, nb, do = job
                        got = 0
                        while got < nb:
                            k = os.preadv(fd, [memoryview(flat[do + got:do + nb])], fo + got)
                            if k <= 0:
                                raise IOError("short read in %s" % path)
                            got += k
                    with ThreadPoolExecutor(max(1, min(threads, len(jobs)))) as ex:
                        list(ex.map(rd, jobs))
                finally:
                    os.close(fd)
            else:                                       # filtered / real h5py object: decode
                crd = pinned_array((self.nbead, self.nstruct, 3), np.float32)
                crd[...] = np.asarray(ds[:], dtype=np.float32)
            self.coordinates = crd
            self.radii = np.asarray(f["radii"][:], dtype=np.float32)
            self.chrom = np.asarray(f["index"]["chrom"][:], dtype=np.int32)
            ci = f["index"]["copy_index"][()]
            if isinstance(ci, np.ndarray):
                ci = ci.tobytes() if ci.dtype.kind in "SV" else ci.item()
            if isinstance(ci, (bytes, np.bytes_)):
                ci = bytes(ci).rstrip(b"\x00").decode("utf-8")
            self.copy_index = CopyIndex.from_dict(json.loads(ci))


_MODES = {"LB": MODE_LB, "GP": MODE_GP, MODE_LB: MODE_LB, MODE_GP: MODE_GP, "lb": MODE_LB, "gp": MODE_GP}


class ActdistEngine:
    def __init__(self, pop: Optional[Population] = None, device: int = 0, *,
                 nbead: Optional[int] = None, nstruct: Optional[int] = None):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        if pop is not None:
            nbead, nstruct = pop.nbead, pop.nstruct
        check(self._lib.igmk_create(int(device), int(nbead), int(nstruct), C.byref(self._ctx)))
        self.device = int(device)
        self.nbead, self.nstruct = int(nbead), int(nstruct)
        self.n_hap = 0
        self._ncopies = None
        self._chrom_hap = None
        self._pending_xyz = None           # host coordinates not yet in HBM (from_hss(upload=False))
        if pop is not None:
            self.upload_coordinates(pop.coordinates)
            self.set_index(pop.copy_index.ptr, pop.copy_index.beads, pop.chrom_hap(), pop.radii)
            self.set_bead_chrom(pop.chrom)

    # -- lifetime --------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.igmk_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @classmethod
    def from_hss(cls, path: str, device: int = 0, staged: "Optional[StagedHss]" = None,
                 upload: bool = True) -> "ActdistEngine":
        """Engine for the population of a .hss file: replaces HssFile(...) + per-pair
        get_bead_crd chunk reads (igm/steps/ActivationDistanceStep.py:202,415-416;
        igm/core/step.py:386-392).  The coordinates go file -> pinned host memory (parallel
        pread of the stored extents, no decoding) -> HBM (double-buffered asynchronous copies
        + re-layout kernel); ``staged`` shares one host copy between the engines of several
        devices.  ``upload=False`` leaves the coordinates in host memory until the first
        ``actdist_buffers`` call, which stages them inside the same pipelined device operation
        (igmk_actdist_host_population); ``stage_pending`` uploads them on their own."""
        st = staged if staged is not None else StagedHss(path)
        eng = cls(nbead=st.nbead, nstruct=st.nstruct, device=device)
        try:
            if upload:
                eng.upload_coordinates(st.coordinates)
            else:
                eng._pending_xyz = st.coordinates
            eng.set_index(st.copy_index.ptr, st.copy_index.beads,
                          np.ascontiguousarray(st.chrom[:len(st.copy_index)]), st.radii)
            eng.set_bead_chrom(st.chrom)
            eng.copy_index, eng.chrom = st.copy_index, st.chrom
        except Exception:
            eng.close()
            raise
        return eng

    # -- staging ---------------------------------------------------------
    def stage_pending(self) -> None:
        """Upload the coordinates ``from_hss(upload=False)`` left in host memory."""
        xyz, self._pending_xyz = self._pending_xyz, None
        if xyz is not None:
            self.upload_coordinates(xyz)

    def reserve_pairs(self, n_pairs: int) -> None:
        """Size the device buffers of the host entry points for lists of up to ``n_pairs``."""
        if n_pairs > getattr(self, "_reserved_pairs", 0):
            check(self._lib.igmk_reserve_pairs(self._ctx, int(n_pairs)))
            self._reserved_pairs = int(n_pairs)

    def actdist_buffers(self, n: int, b_i, b_j, b_pwish, b_plast, b_out, contact_range: float = 2.0,
                        it_corr: int = 0, mode="LB", algo: int = ALGO_FAST) -> None:
        """A-step on caller-owned (ideally page-locked) host buffers, no copies on the Python
        side.  A population left pending by ``from_hss(upload=False)`` is staged by the same
        call, its upload hidden behind the pair kernels."""
        a = (int(n), ptr(b_i), ptr(b_j), ptr(b_pwish), ptr(b_plast), float(np.float32(contact_range)),
             int(it_corr), _MODES[mode], int(algo), ptr(b_out))
        xyz = self._pending_xyz
        if xyz is not None and n > 0:
            check(self._lib.igmk_actdist_host_population(self._ctx, ptr(xyz), *a))
            self._pending_xyz = None
        else:
            self.stage_pending()
            check(self._lib.igmk_actdist_host(self._ctx, *a))

    def upload_coordinates(self, xyz, bead0: int = 0) -> None:
        """xyz: (nb, nstruct, 3) float32, bead-major - NumPy array (host) or a
        CUDA torch tensor on this engine's device."""
        on_device = 0
        if isinstance(xyz, np.ndarray):
            xyz = np.ascontiguousarray(xyz, dtype=np.float32)
            shape = xyz.shape
        else:
            if not xyz.is_cuda:
                xyz = xyz.contiguous().numpy()
                return self.upload_coordinates(xyz, bead0)
            if str(xyz.dtype) != "torch.float32":
                raise TypeError("coordinates must be float32")
            xyz = xyz.contiguous()
            shape = tuple(xyz.shape)
            on_device = 1
        if len(shape) != 3 or shape[1] != self.nstruct or shape[2] != 3:
            raise ValueError("coordinates must be (nbead, %d, 3), got %r" % (self.nstruct, shape))
        check(self._lib.igmk_upload_coords_range(self._ctx, ptr(xyz), int(bead0), int(shape[0]), on_device))

    def copy_coordinates_from(self, other: "ActdistEngine", bead0: int, nb: int) -> None:
        """Rows [bead0, bead0 + nb) from another engine of this process (device to device,
        NVLink between peers)."""
        check(self._lib.igmk_copy_coords_peer(self._ctx, other._ctx, int(bead0), int(nb)))

    def coords_tensor(self):
        """The staged population as a torch tensor (rows, 3 * npad) float32 on this engine's
        device, zero-copy (rows = nbead + 1 + 64 spare) - for the caller's collectives."""
        import torch
        p, rows, rf = C.c_void_p(), C.c_int64(), C.c_int64()
        check(self._lib.igmk_coords_device(self._ctx, C.byref(p), C.byref(rows), C.byref(rf)))

        class _Buf:
            __cuda_array_interface__ = {"shape": (int(rows.value), int(rf.value)), "typestr": "<f4",
                                        "data": (int(p.value), False), "version": 2}
        return torch.as_tensor(_Buf(), device=torch.device("cuda", self.device))

    def set_index(self, copy_ptr, copy_beads, chrom_hap, radii) -> None:
        copy_ptr = np.ascontiguousarray(copy_ptr, dtype=np.int32)
        copy_beads = np.ascontiguousarray(copy_beads, dtype=np.int32)
        chrom_hap = np.ascontiguousarray(chrom_hap, dtype=np.int32)
        radii = np.ascontiguousarray(radii, dtype=np.float32)
        n_hap = len(copy_ptr) - 1
        if len(chrom_hap) != n_hap or len(radii) != self.nbead:
            raise ValueError("index arrays have inconsistent lengths")
        check(self._lib.igmk_set_index(self._ctx, n_hap, ptr(copy_ptr), ptr(copy_beads),
                                       ptr(chrom_hap), ptr(radii)))
        self.n_hap = n_hap
        self._ncopies = np.diff(copy_ptr)
        self._chrom_hap = chrom_hap
        # whole-chromosome ploidy (igm/_preprocess.py:66-86): every locus of a chromosome
        # has the same number of copies, so no intra-chromosomal pair can violate q5
        nchrom = int(chrom_hap.max()) + 1 if n_hap else 0
        lo = np.full(nchrom, np.iinfo(np.int32).max, np.int64)
        hi = np.full(nchrom, -1, np.int64)
        np.minimum.at(lo, chrom_hap, self._ncopies)
        np.maximum.at(hi, chrom_hap, self._ncopies)
        self._uniform_ploidy = bool(np.all((lo == hi) | (hi < 0))) and (n_hap == 0 or int(chrom_hap.min()) >= 0)

    # -- A-step ----------------------------------------------------------
    def _validate_pairs(self, i, j, mode):
        if len(i) == 0:
            return
        if i.min() < 0 or j.min() < 0 or i.max() >= self.n_hap or j.max() >= self.n_hap:
            raise ValueError("pair index out of range [0, %d)" % self.n_hap)
        if mode == MODE_LB and not getattr(self, "_uniform_ploidy", False):
            # quirk q5: the reference's intra branch reads uninitialised
            # memory when the two loci have different copy counts - refuse instead.
            bad = (self._chrom_hap[i] == self._chrom_hap[j]) & (self._ncopies[i] != self._ncopies[j]) & (i != j)
            if bad.any():
                k = int(np.nonzero(bad)[0][0])
                raise ValueError("intra-chromosomal pair (%d, %d) has unequal copy counts" % (i[k], j[k]))

    def actdist(self, i, j, pwish, plast=None, contact_range: float = 2.0, it_corr: int = 0,
                mode="LB", algo: int = ALGO_FAST) -> np.ndarray:
        """Host arrays in, structured result array (PAIR_RESULT_DTYPE) out.
        One entry per pair, input order.  Copies are inside the call."""
        mode = _MODES[mode]
        i = np.ascontiguousarray(i, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        pwish = np.ascontiguousarray(pwish, dtype=np.float64)
        plast = np.zeros(len(i), np.float64) if plast is None else np.ascontiguousarray(plast, dtype=np.float64)
        if not (len(i) == len(j) == len(pwish) == len(plast)):
            raise ValueError("pair arrays must have equal length")
        self._validate_pairs(i, j, mode)
        out = np.zeros(len(i), dtype=PAIR_RESULT_DTYPE)
        check(self._lib.igmk_actdist_host(self._ctx, len(i), ptr(i), ptr(j), ptr(pwish), ptr(plast),
                                          float(np.float32(contact_range)), int(it_corr), mode, int(algo),
                                          ptr(out)))
        return out

    def actdist_with_population(self, xyz, i, j, pwish, plast=None, contact_range: float = 2.0,
                                it_corr: int = 0, mode="LB", algo: int = ALGO_FAST, out=None) -> np.ndarray:
        """``upload_coordinates(xyz)`` + ``actdist(...)`` as ONE pipelined device operation
        (igmk_actdist_host_population): the population upload hides behind the pair kernels.
        ``xyz``: host array (nbead, nstruct, 3) float32, ideally page-locked
        (``pinned_array`` / ``StagedHss.coordinates``); the index must be set."""
        mode = _MODES[mode]
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        if xyz.shape != (self.nbead, self.nstruct, 3):
            raise ValueError("coordinates must be (%d, %d, 3), got %r" % (self.nbead, self.nstruct, xyz.shape))
        i = np.ascontiguousarray(i, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        pwish = np.ascontiguousarray(pwish, dtype=np.float64)
        plast = np.zeros(len(i), np.float64) if plast is None else np.ascontiguousarray(plast, dtype=np.float64)
        if not (len(i) == len(j) == len(pwish) == len(plast)):
            raise ValueError("pair arrays must have equal length")
        self._validate_pairs(i, j, mode)
        if out is None:
            out = np.zeros(len(i), dtype=PAIR_RESULT_DTYPE)
        check(self._lib.igmk_actdist_host_population(self._ctx, ptr(xyz), len(i), ptr(i), ptr(j), ptr(pwish),
                                                     ptr(plast), float(np.float32(contact_range)), int(it_corr),
                                                     mode, int(algo), ptr(out)))
        return out

    def actdist_device(self, d_i, d_j, d_pwish, d_plast, d_out, n_pairs: Optional[int] = None,
                       contact_range: float = 2.0, it_corr: int = 0, mode="LB",
                       algo: int = ALGO_FAST, stream: int = 0) -> None:
        """Device buffers (torch CUDA tensors or raw addresses); asynchronous."""
        n = int(n_pairs if n_pairs is not None else d_i.numel())
        check(self._lib.igmk_actdist_device(self._ctx, n, ptr(d_i), ptr(d_j), ptr(d_pwish),
                                            ptr(d_plast), float(np.float32(contact_range)), int(it_corr),
                                            _MODES[mode], int(algo), ptr(d_out), stream or None))

    def actdist_device_peers(self, d_i, d_j, d_pwish, d_plast, d_peer_slices, n_peers: int,
                             n_pairs: Optional[int] = None, contact_range: float = 2.0,
                             it_corr: int = 0, mode="LB", stream: int = 0) -> None:
        """Multi-GPU form: raw results are stored into every GPU's gather buffer from
        inside the kernel (d_peer_slices: device int64 array of n_peers addresses);
        finish_results() must follow a cross-rank barrier."""
        n = int(n_pairs if n_pairs is not None else d_i.numel())
        check(self._lib.igmk_actdist_device_peers(self._ctx, n, ptr(d_i), ptr(d_j), ptr(d_pwish),
                                                  ptr(d_plast), float(np.float32(contact_range)),
                                                  int(it_corr), _MODES[mode], ptr(d_peer_slices),
                                                  int(n_peers), stream or None))

    def sel_flat_idx(self, i, j, res, mode="LB") -> np.ndarray:
        """Index of the selected element of every pair in ``d_sq[0:npc].ravel()``
        (row * nstruct + structure; lowest index among ties; -1 without a record)."""
        i = np.ascontiguousarray(i, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        res = np.ascontiguousarray(res, dtype=PAIR_RESULT_DTYPE)
        out = np.full(len(i), -1, dtype=np.int32)
        check(self._lib.igmk_actdist_sel_index_host(self._ctx, len(i), ptr(i), ptr(j), ptr(res), _MODES[mode], ptr(out)))
        return out

    def last_redo_count(self) -> int:
        """Pairs of the most recent A-step launch that the list-form kernel handed back
        to the key-array kernels (diagnostic; synchronises the device)."""
        return int(self._lib.igmk_last_redo_count(self._ctx))

    def finish_results(self, d_results, n: int, stream: int = 0) -> None:
        check(self._lib.igmk_finish_results_device(self._ctx, ptr(d_results), int(n), stream or None))

    # -- M-step restraint selection (K3) ----------------------------------
    KINDS = {"intra": 0, "inter": 1, "any": 2}

    def set_bead_chrom(self, chrom_bead) -> None:
        chrom_bead = np.ascontiguousarray(chrom_bead, dtype=np.int32)
        if len(chrom_bead) != self.nbead:
            raise ValueError("chrom must have one entry per bead")
        check(self._lib.igmk_set_bead_chrom(self._ctx, ptr(chrom_bead)))

    def restraint_select(self, row, col, dist, kind="intra"):
        """(bitmap, counts): bitmap[k, s // 32] >> (s % 32) & 1 says whether structure s
        gets the restraint of actdist record k (intraHiC / interHiC._apply)."""
        row = np.ascontiguousarray(row, dtype=np.int32)
        col = np.ascontiguousarray(col, dtype=np.int32)
        dist = np.ascontiguousarray(dist, dtype=np.float32)
        if not (len(row) == len(col) == len(dist)):
            raise ValueError("record arrays must have equal length")
        words = int(self._lib.igmk_restraint_words(self._ctx))
        bitmap = np.zeros((len(row), words), dtype=np.uint32)
        counts = np.zeros(len(row), dtype=np.int32)
        check(self._lib.igmk_restraint_select_host(self._ctx, len(row), ptr(row), ptr(col), ptr(dist),
                                                   self.KINDS[kind], ptr(bitmap), ptr(counts)))
        return bitmap, counts

    @staticmethod
    def records_of_structure(bitmap: np.ndarray, s: int) -> np.ndarray:
        """Indices of the records whose restraint structure s receives, in record order
        (what _apply iterates for one model)."""
        return np.nonzero((bitmap[:, s >> 5] >> np.uint32(s & 31)) & np.uint32(1))[0]

    # -- rank matching (K5: FISH / polymer assignment) ---------------------
    def rank_match(self, a, b=None, reduce="min", target=None, want_rank=True, want_value=True):
        """Per item the reduced (min / max over copy combinations) float32 distance of every
        structure, its rank in the population (ties by structure index) and, with
        ``target`` (one sorted row of nstruct values, or one row per item), the target handed
        to each structure: get_min_max_and_idx + ``target[idx]`` of
        igm/steps/FishAssignmentStep.py:60-79,189-193 and get_polymer_dists of
        igm/steps/PolymerAssignmentStep.py:24-33.  ``a`` / ``b``: (n_items, 2) bead ids, second
        column -1 for single-copy loci; ``b=None``: radial distances.  Returns a dict with
        the requested arrays ``value``, ``rank``, ``matched`` of shape (n_items, nstruct)."""
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, 2)
        n = len(a)
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, 2)
            if len(b) != n:
                raise ValueError("a and b must have one row per item")
        stride = 0
        if target is not None:
            target = np.ascontiguousarray(target, dtype=np.float32)
            if target.ndim == 2:
                if target.shape != (n, self.nstruct):
                    raise ValueError("target must be (n_items, nstruct) or (nstruct,)")
                stride = self.nstruct
            elif target.shape != (self.nstruct,):
                raise ValueError("target must be (n_items, nstruct) or (nstruct,)")
        out = {}
        if target is not None:
            out["matched"] = np.zeros((n, self.nstruct), np.float32)
        if want_rank:
            out["rank"] = np.zeros((n, self.nstruct), np.int32)
        if want_value:
            out["value"] = np.zeros((n, self.nstruct), np.float32)
        check(self._lib.igmk_rank_match_host(
            self._ctx, n, ptr(a), ptr(b) if b is not None else None, {"min": 0, "max": 1}[reduce],
            ptr(target) if target is not None else None, stride,
            ptr(out["matched"]) if "matched" in out else None,
            ptr(out["rank"]) if "rank" in out else None,
            ptr(out["value"]) if "value" in out else None))
        return out

    # -- SPRITE (K4) ------------------------------------------------------
    def sprite_rg2(self, clusters):
        """clusters: list of clusters, each a list of regions, each a list of bead ids
        (the alternative locations).  Returns a list of (rg2s[nstruct], best_structure,
        copy_idxs[nstruct, n_regions]) - get_rgs2's return value per cluster
        (igm/cython_compiled/sprite.pyx:36-101)."""
        region_ptr, copy_ptr, beads = [0], [0], []
        for cl in clusters:
            for reg in cl:
                beads.extend(int(b) for b in reg)
                copy_ptr.append(len(beads))
            region_ptr.append(len(copy_ptr) - 1)
        region_ptr = np.asarray(region_ptr, np.int32)
        copy_ptr = np.asarray(copy_ptr, np.int32)
        beads = np.asarray(beads, np.int32)
        n = len(clusters)
        rg2s = np.zeros((n, self.nstruct), np.float32)
        cidx = np.zeros(int(region_ptr[-1]) * self.nstruct, np.int32)
        ms = np.zeros(n, np.int32)
        check(self._lib.igmk_sprite_rg2_host(self._ctx, n, ptr(region_ptr), ptr(copy_ptr), ptr(beads),
                                             ptr(rg2s), ptr(cidx), ptr(ms)))
        out = []
        for k in range(n):
            m = int(region_ptr[k + 1] - region_ptr[k])
            lo = int(region_ptr[k]) * self.nstruct
            out.append((rg2s[k], int(ms[k]), cidx[lo:lo + m * self.nstruct].reshape(self.nstruct, m)))
        return out

    def sprite_cluster_rg2(self, clusters):
        """Rg^2 of whole clusters under a per-structure choice of copies (K4b) - the second
        half of compute_gyration_radius (igm/cython_compiled/sprite.pyx:238-283).

        clusters: list of ``(segments, groups, sel)``: ``segments`` a list of bead-id lists
        (the copies of each segment, in the reference's concatenation order), ``groups`` the
        selection column each segment follows, ``sel`` an int array (nstruct, n_columns) of
        chosen copies (get_rgs2's copy_idxs; negative = from the end).  Returns a float32
        array (n_clusters, nstruct)."""
        seg_ptr, loc_ptr, beads, seg_group, group_ptr, sels = [0], [0], [], [], [0], []
        for segments, groups, sel in clusters:
            sel = np.ascontiguousarray(sel, dtype=np.int32)
            if sel.ndim != 2 or sel.shape[0] != self.nstruct:
                raise ValueError("sel must have shape (nstruct, n_columns)")
            if len(groups) != len(segments):
                raise ValueError("one selection column per segment is required")
            for reg in segments:
                beads.extend(int(b) for b in reg)
                loc_ptr.append(len(beads))
            seg_group.extend(int(g) for g in groups)
            seg_ptr.append(len(loc_ptr) - 1)
            group_ptr.append(group_ptr[-1] + sel.shape[1])
            sels.append(sel.reshape(-1))
        n = len(clusters)
        rg2s = np.zeros((n, self.nstruct), np.float32)
        if n == 0:
            return rg2s
        a = [np.asarray(x, np.int32) for x in (seg_ptr, loc_ptr, beads, seg_group, group_ptr)]
        sel_all = np.ascontiguousarray(np.concatenate(sels), dtype=np.int32)
        check(self._lib.igmk_sprite_cluster_rg2_host(self._ctx, n, ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(a[3]),
                                                     ptr(a[4]), ptr(sel_all), ptr(rg2s)))
        return rg2s

    # -- DamID ----------------------------------------------------------
    def damid_actdist(self, loci, p_exp, plast=None, nucleus_radius: float = 5000.0,
                      contact_range: float = 0.05, it_corr: int = 0) -> np.ndarray:
        """get_damid_actdist_I for every locus (spherical envelope); one result per locus."""
        loci = np.ascontiguousarray(loci, dtype=np.int32)
        p_exp = np.ascontiguousarray(p_exp, dtype=np.float32)
        plast = np.zeros(len(loci), np.float32) if plast is None else np.ascontiguousarray(plast, dtype=np.float32)
        if not (len(loci) == len(p_exp) == len(plast)):
            raise ValueError("locus arrays must have equal length")
        if len(loci) and (loci.min() < 0 or loci.max() >= self.n_hap):
            raise ValueError("locus index out of range [0, %d)" % self.n_hap)
        out = np.zeros(len(loci), dtype=PAIR_RESULT_DTYPE)
        check(self._lib.igmk_damid_actdist_host(self._ctx, len(loci), ptr(loci), ptr(p_exp), ptr(plast),
                                                float(nucleus_radius), float(contact_range), int(it_corr),
                                                ptr(out)))
        return out

    def expand_damid_records(self, loci, res, copy_ptr, copy_beads):
        """(loc, dist, prob) columns of damid_actdist.hdf5: one record per copy of each
        locus, in locus order (DamidActivationDistanceStep.py:468, :287-296)."""
        loci = np.asarray(loci, dtype=np.int64)
        nrec = res["nrec"].astype(np.int64)
        first = np.asarray(copy_ptr, dtype=np.int64)[loci]
        assert np.array_equal(nrec, np.asarray(copy_ptr, dtype=np.int64)[loci + 1] - first)
        rep = np.repeat(np.arange(len(loci)), nrec)
        within = np.arange(len(rep)) - np.repeat(np.cumsum(nrec) - nrec, nrec)
        loc = np.asarray(copy_beads)[first[rep] + within].astype(np.int32)
        return loc, res["dist"][rep].astype(np.float32), res["prob"][rep].astype(np.float32)

    def expand_records(self, i, j, res) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """Pair results -> (row, col, dist, prob), the four actdist.hdf5 columns."""
        i = np.ascontiguousarray(i, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        res = np.ascontiguousarray(res, dtype=PAIR_RESULT_DTYPE)
        cap = int(res["nrec"].sum())
        row = np.empty(cap, np.int32)
        col = np.empty(cap, np.int32)
        dist = np.empty(cap, np.float32)
        prob = np.empty(cap, np.float32)
        n = C.c_int64(0)
        check(self._lib.igmk_expand_records(self._ctx, len(i), ptr(i), ptr(j), ptr(res), ptr(row),
                                            ptr(col), ptr(dist), ptr(prob), cap, C.byref(n)))
        assert n.value == cap
        return row, col, dist, prob

    # -- contact frequency -------------------------------------------------
    def contact_counts(self, row0: int, nrows: int, col0: int, ncols: int,
                       contact_range: float = 2.0, strict: bool = False) -> np.ndarray:
        out = np.zeros((nrows, ncols), dtype=np.uint32)
        check(self._lib.igmk_contact_counts_host(self._ctx, int(row0), int(nrows), int(col0),
                                                 int(ncols), float(np.float32(contact_range)),
                                                 1 if strict else 0, ptr(out)))
        return out

    def contact_counts_device(self, row0, nrows, col0, ncols, d_counts, contact_range=2.0,
                              strict=False, stream: int = 0) -> None:
        check(self._lib.igmk_contact_counts_device(self._ctx, int(row0), int(nrows), int(col0),
                                                   int(ncols), float(np.float32(contact_range)),
                                                   1 if strict else 0, ptr(d_counts), stream or None))

    def contact_counts_haploid(self, row0: int, nrows: int, col0: int, ncols: int,
                               contact_range: float = 2.0, strict: bool = False) -> np.ndarray:
        """Copy-summed counts for a tile of haploid loci (sumCopies of the bead map)."""
        out = np.zeros((nrows, ncols), dtype=np.uint32)
        check(self._lib.igmk_contact_counts_haploid_host(self._ctx, int(row0), int(nrows), int(col0),
                                                         int(ncols), float(np.float32(contact_range)),
                                                         1 if strict else 0, ptr(out)))
        return out

    def contact_counts_haploid_device(self, row0, nrows, col0, ncols, d_counts, contact_range=2.0,
                                      strict=False, stream: int = 0) -> None:
        check(self._lib.igmk_contact_counts_haploid_device(self._ctx, int(row0), int(nrows), int(col0),
                                                           int(ncols), float(np.float32(contact_range)),
                                                           1 if strict else 0, ptr(d_counts), stream or None))

    def last_kernel_ms(self) -> float:
        return float(self._lib.igmk_last_kernel_ms(self._ctx))


def launch_count() -> int:
    return int(_lib.load().igmk_launch_count())
