"""TEST INFRASTRUCTURE ONLY - oracles for K4 (SPRITE cluster radius of gyration).

1. ``ref_get_rgs2``: the reference's OWN native code - get_rg2s_cpp
   (igm/cython_compiled/cpp_sprite_assignment.cpp:79-143) compiled unmodified from
   /root/reference by oracle/Makefile into oracle/_ref/libsprite_ref.so and called with
   the argument conventions of its Cython wrapper get_rgs2 (sprite.pyx:36-101).
2. ``get_rgs2_port``: a NumPy float32 restatement, pinned against (1) and against the
   three known answers of igm/cython_compiled/tests.py (executed values; the comment of
   case 2 there is a stale copy of case 1).

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libsprite_ref.so")
INF = np.float32(100000000.0)       # cpp_sprite_assignment.cpp:4
_lib = None


def ref_available():
    return os.path.exists(REF_LIB)


def ref_get_rgs2(crds, copies_num):
    """crds: (B, N, 3) float32, copies_num: (M,) int32 -> (rg2s, best_structure, copy_idxs)."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(REF_LIB)
        _lib.ref_get_rg2s.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    crds = np.ascontiguousarray(crds, np.float32)
    copies_num = np.ascontiguousarray(copies_num, np.int32)
    n_bead, n_struct = crds.shape[0], crds.shape[1]
    m = len(copies_num)
    rg2s = np.zeros(n_struct, np.float32)
    cidx = np.zeros((n_struct, m), np.int32)
    best = C.c_int(-1)                      # the reference leaves it unset when nothing is below INF
    _lib.ref_get_rg2s(crds.ctypes.data, n_struct, n_bead, m, copies_num.ctypes.data,
                      rg2s.ctypes.data, cidx.ctypes.data, C.byref(best))
    return rg2s, int(best.value), cidx


def get_rgs2_port(crds, copies_num):
    """NumPy float32 restatement of get_rg2s_cpp / gyration_radius_sq (:49-61, :79-143)."""
    crds = np.asarray(crds, np.float32)
    copies_num = [int(k) for k in copies_num]
    n_struct, m = crds.shape[1], len(copies_num)
    first = np.concatenate([[0], np.cumsum(copies_num)[:-1]]).astype(int)
    ncomb = int(np.prod(copies_num))
    f32 = np.float32
    fn = f32(m)
    best = np.full(n_struct, INF, np.float32)
    best_k = np.full(n_struct, -1, np.int64)
    for k in range(ncomb):
        kk, sel = k, []
        for i in range(m):
            sel.append(first[i] + kk % copies_num[i])
            kk //= copies_num[i]
        pts = crds[sel]                                   # (m, N, 3)
        mean = np.zeros((n_struct, 3), np.float32)
        for i in range(m):
            mean = (mean + pts[i]).astype(np.float32)
        mean = (mean / fn).astype(np.float32)
        rg = np.zeros(n_struct, np.float32)
        for i in range(m):
            d = (pts[i] - mean).astype(np.float32)
            sq = (d * d).astype(np.float32)
            rg = (rg + ((sq[:, 0] + sq[:, 1]).astype(np.float32) + sq[:, 2]).astype(np.float32)).astype(np.float32)
        rg2 = (rg / fn).astype(np.float32)
        upd = rg2 < best
        best[upd] = rg2[upd]
        best_k[upd] = k
    cidx = np.full((n_struct, m), -1, np.int32)
    kk = best_k.copy()
    for i in range(m):
        cidx[:, i] = np.where(best_k >= 0, kk % copies_num[i], -1)
        kk = np.where(best_k >= 0, kk // copies_num[i], kk)
    below = np.nonzero(best < INF)[0]
    best_struct = int(below[np.argmin(best[below])]) if len(below) else -1
    return best, best_struct, cidx
