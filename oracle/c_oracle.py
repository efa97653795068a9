"""TEST INFRASTRUCTURE ONLY - ctypes wrapper of oracle/actdist_oracle.c
(plain-C restatement of the reference A-step; `make -C oracle`)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libactdist_oracle.so")
_lib = None


def available():
    return os.path.exists(LIB)


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        vp = C.c_void_p
        _lib.actdist_oracle_pairs.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_long, vp, vp, vp, vp,
                                             C.c_float, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int]
        _lib.contact_oracle_counts.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_float,
                                               C.c_int, vp, C.c_int]
    return _lib


def run_pairs(ii, jj, pwish, plast, coords, radii, chrom_hap, copy_ptr, copy_beads, it_corr,
              contact_range=2.0, mode=0, nthreads=0):
    """Same outputs as actdist_oracle.details_to_arrays(run_pairs(...))."""
    lib = load()
    ii = np.ascontiguousarray(ii, np.int32)
    jj = np.ascontiguousarray(jj, np.int32)
    pwish = np.ascontiguousarray(pwish, np.float64)
    plast = np.ascontiguousarray(plast, np.float64)
    coords = np.ascontiguousarray(coords, np.float32)
    radii = np.ascontiguousarray(radii, np.float32)
    chrom_hap = np.ascontiguousarray(chrom_hap, np.int32)
    copy_ptr = np.ascontiguousarray(copy_ptr, np.int32)
    copy_beads = np.ascontiguousarray(copy_beads, np.int32)
    n = len(ii)
    out = {"d2_sel_bits": np.zeros(n, np.uint32), "contact_count": np.zeros(n, np.int32),
           "o": np.zeros(n, np.int32), "p": np.zeros(n, np.float64), "nrec": np.zeros(n, np.int32)}
    rc = lib.actdist_oracle_pairs(coords.ctypes.data, coords.shape[1], radii.ctypes.data,
                                  chrom_hap.ctypes.data, copy_ptr.ctypes.data, copy_beads.ctypes.data,
                                  n, ii.ctypes.data, jj.ctypes.data, pwish.ctypes.data, plast.ctypes.data,
                                  float(np.float32(contact_range)), int(it_corr), int(mode),
                                  out["d2_sel_bits"].ctypes.data, out["contact_count"].ctypes.data,
                                  out["o"].ctypes.data, out["p"].ctypes.data, out["nrec"].ctypes.data,
                                  int(nthreads))
    if rc:
        raise MemoryError("actdist_oracle_pairs")
    return out


def contact_counts(coords, radii, rows, cols, contact_range=2.0, strict=False, nthreads=0):
    lib = load()
    coords = np.ascontiguousarray(coords, np.float32)
    radii = np.ascontiguousarray(radii, np.float32)
    rows = np.ascontiguousarray(rows, np.int32)
    cols = np.ascontiguousarray(cols, np.int32)
    out = np.zeros((len(rows), len(cols)), np.uint32)
    lib.contact_oracle_counts(coords.ctypes.data, coords.shape[1], radii.ctypes.data, len(rows),
                              rows.ctypes.data, len(cols), cols.ctypes.data,
                              float(np.float32(contact_range)), 1 if strict else 0, out.ctypes.data,
                              int(nthreads))
    return out
