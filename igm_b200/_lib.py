"""ctypes binding of libigmk.so (C ABI: include/igmk.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (or
``python -m igm_b200.build``).  There is no CPU fallback: if the library is
missing, importing this module still works (so CPU-only tooling can inspect the
package) but any attempt to use it raises ``IgmkError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IGMK_LIB_PATH") or os.path.join(HERE, "libigmk.so")

IGMK_OK, IGMK_EINVAL, IGMK_ECUDA, IGMK_ESTATE, IGMK_ELIMIT = 0, -1, -2, -3, -4
MODE_LB, MODE_GP = 0, 1
ALGO_FAST, ALGO_SIMPLE = 0, 1

# mirrors `struct igmk_pair_result` (32 bytes)
PAIR_RESULT_DTYPE = np.dtype([
    ("d2_sel_bits", np.uint32), ("contact_count", np.int32), ("o", np.int32),
    ("nrec", np.int32), ("p", np.float64), ("dist", np.float32), ("prob", np.float32)],
    align=True)
assert PAIR_RESULT_DTYPE.itemsize == 32

# every symbol include/igmk.h declares
EXPORTS = (
    "igmk_version", "igmk_last_error", "igmk_launch_count", "igmk_create", "igmk_destroy",
    "igmk_upload_coords", "igmk_upload_coords_range", "igmk_set_index",
    "igmk_actdist_device", "igmk_actdist_host", "igmk_actdist_device_peers",
    "igmk_finish_results_device", "igmk_expand_records", "igmk_filter_candidates", "igmk_join_plast",
    "igmk_damid_actdist_device", "igmk_damid_actdist_host",
    "igmk_contact_counts_device", "igmk_contact_counts_host",
    "igmk_contact_counts_haploid_device", "igmk_contact_counts_haploid_host",
    "igmk_set_bead_chrom", "igmk_restraint_words", "igmk_restraint_select_device",
    "igmk_restraint_select_host", "igmk_sprite_rg2_host", "igmk_sprite_cluster_rg2_host",
    "igmk_rank_match_device", "igmk_rank_match_host",
    "igmk_host_alloc", "igmk_host_free", "igmk_last_kernel_ms", "igmk_last_redo_count",
    "igmk_actdist_sel_index_device", "igmk_actdist_sel_index_host",
    "igmk_actdist_host_population", "igmk_coords_device", "igmk_copy_coords_peer",
    "igmk_reserve_pairs",
)


class IgmkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("igmk error %d: %s" % (code, msg))
        self.code = code


_lib: Optional[C.CDLL] = None


def _declare(lib: C.CDLL) -> None:
    vp, i32p, f32p, f64p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.igmk_version.restype = C.c_int
    lib.igmk_last_error.restype = C.c_char_p
    lib.igmk_launch_count.restype = C.c_int64
    lib.igmk_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.igmk_destroy.argtypes = [vp]
    lib.igmk_upload_coords.argtypes = [vp, f32p, C.c_int]
    lib.igmk_upload_coords_range.argtypes = [vp, f32p, C.c_int, C.c_int, C.c_int]
    lib.igmk_coords_device.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.igmk_copy_coords_peer.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.igmk_reserve_pairs.argtypes = [vp, C.c_int64]
    lib.igmk_set_index.argtypes = [vp, C.c_int, i32p, i32p, i32p, f32p]
    lib.igmk_actdist_device.argtypes = [vp, C.c_int64, i32p, i32p, f64p, f64p, C.c_float,
                                        C.c_int, C.c_int, C.c_int, vp, vp]
    lib.igmk_actdist_host.argtypes = [vp, C.c_int64, i32p, i32p, f64p, f64p, C.c_float,
                                      C.c_int, C.c_int, C.c_int, vp]
    lib.igmk_actdist_host_population.argtypes = [vp, f32p, C.c_int64, i32p, i32p, f64p, f64p, C.c_float,
                                                 C.c_int, C.c_int, C.c_int, vp]
    lib.igmk_actdist_device_peers.argtypes = [vp, C.c_int64, i32p, i32p, f64p, f64p, C.c_float,
                                              C.c_int, C.c_int, vp, C.c_int, vp]
    lib.igmk_finish_results_device.argtypes = [vp, vp, C.c_int64, vp]
    lib.igmk_damid_actdist_device.argtypes = [vp, C.c_int64, i32p, f32p, f32p, C.c_double, C.c_double,
                                              C.c_int, vp, vp]
    lib.igmk_damid_actdist_host.argtypes = [vp, C.c_int64, i32p, f32p, f32p, C.c_double, C.c_double,
                                            C.c_int, vp]
    lib.igmk_expand_records.argtypes = [vp, C.c_int64, i32p, i32p, vp, i32p, i32p, f32p, f32p,
                                        C.c_int64, C.POINTER(C.c_int64)]
    lib.igmk_contact_counts_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                               C.c_int, vp, vp]
    lib.igmk_contact_counts_host.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                             C.c_int, vp]
    lib.igmk_contact_counts_haploid_device.argtypes = lib.igmk_contact_counts_device.argtypes
    lib.igmk_contact_counts_haploid_host.argtypes = lib.igmk_contact_counts_host.argtypes
    lib.igmk_set_bead_chrom.argtypes = [vp, i32p]
    lib.igmk_restraint_words.argtypes = [vp]
    lib.igmk_restraint_select_device.argtypes = [vp, C.c_int64, i32p, i32p, f32p, C.c_int, vp, vp, vp]
    lib.igmk_restraint_select_host.argtypes = [vp, C.c_int64, i32p, i32p, f32p, C.c_int, vp, vp]
    lib.igmk_filter_candidates.restype = C.c_int64
    lib.igmk_filter_candidates.argtypes = [C.c_int64, vp, i32p, f32p, i32p, C.c_int, C.c_float, C.c_int, C.c_float,
                                           i32p, i32p, f64p, C.c_int64]
    lib.igmk_join_plast.argtypes = [C.c_int64, i32p, i32p, f32p, C.c_int32, C.c_int64, i32p, i32p, f64p]
    lib.igmk_sprite_rg2_host.argtypes = [vp, C.c_int, i32p, i32p, i32p, f32p, i32p, i32p]
    lib.igmk_sprite_cluster_rg2_host.argtypes = [vp, C.c_int, i32p, i32p, i32p, i32p, i32p, i32p, f32p]
    lib.igmk_rank_match_device.argtypes = [vp, C.c_int64, i32p, i32p, C.c_int, f32p, C.c_int64, vp, vp, vp, vp]
    lib.igmk_rank_match_host.argtypes = [vp, C.c_int64, i32p, i32p, C.c_int, f32p, C.c_int64, vp, vp, vp]
    lib.igmk_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_int64]
    lib.igmk_host_free.argtypes = [vp]
    lib.igmk_last_kernel_ms.argtypes = [vp]
    lib.igmk_last_kernel_ms.restype = C.c_float
    lib.igmk_last_redo_count.argtypes = [vp]
    lib.igmk_actdist_sel_index_device.argtypes = [vp, C.c_int64, i32p, i32p, vp, C.c_int, i32p, vp]
    lib.igmk_actdist_sel_index_host.argtypes = [vp, C.c_int64, i32p, i32p, vp, C.c_int, i32p]
    lib.igmk_last_redo_count.restype = C.c_int64
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("igmk_version",):
            fn.restype = C.c_int


def load() -> C.CDLL:
    """Load libigmk.so; raises IgmkError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IgmkError(IGMK_ESTATE, "%s not found - run `python __graft_entry__.py` (build()) "
                            "first; igm_b200 has no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        _declare(lib)
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != IGMK_OK:
        raise IgmkError(rc, load().igmk_last_error().decode("utf-8", "replace"))


def ptr(a) -> int:
    """Raw address of a C-contiguous NumPy array or a torch tensor, or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError("cannot take the address of %r" % type(a))
