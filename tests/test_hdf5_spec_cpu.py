"""The HDF5 writer against the file-format specification and against files written by the
real library.

igm_b200/hdf5.py writes actdist.hdf5 without libhdf5; its readers in production are h5py
(ActivationDistanceDB, igm/restraints/intra_hic.py:64-113; the next A-step's setup,
igm/steps/ActivationDistanceStep.py:145-151).  h5py is not installed here, so acceptance by
the real library is argued in two independent steps:
  1. tests/h5spec.py - a structural checker written from the format specification, sharing no
     code with the writer/reader under test - accepts the reference's demo files, which WERE
     written by libhdf5 through h5py (this pins the checker);
  2. the same checker accepts every file the writer produces, and the header messages that
     define how the bytes are interpreted (dataspace, datatype, fill value, layout class /
     version) are byte-identical to those libhdf5 stored for datasets of the same shape and
     type in the demo files.
"""
import os
import struct

import numpy as np
import pytest

from igm_b200 import hdf5
from tests import h5spec
from tests import helpers as H

needs_demo = pytest.mark.skipif(not H.have_demo(), reason="oracle/_ref/demo not present")


@needs_demo
def test_checker_accepts_files_written_by_libhdf5():
    for path, nobj in ((H.DEMO_HSS, 24), (H.DEMO_HCS, 21)):
        objs = h5spec.check(path)
        assert len(objs) == nobj
        kinds = {v["kind"] for v in objs.values()}
        assert kinds == {"group", "dataset"}
    objs = h5spec.check(H.DEMO_HSS)
    assert objs["/coordinates"]["shape"] == (3008, 100, 3) and objs["/coordinates"]["layout_class"] == 2
    assert objs["/index/chromstr"]["layout_class"] == 1          # a contiguous dataset from the real library


def test_checker_rejects_broken_files(tmp_path):
    p = str(tmp_path / "a.h5")
    hdf5.write_h5(p, {"row": np.arange(10, dtype=np.int32)})
    h5spec.check(p)
    raw = bytearray(open(p, "rb").read())
    for off, val, what in ((8, 1, "superblock version"), (13, 4, "offset size")):
        bad = bytearray(raw)
        bad[off] = val
        q = str(tmp_path / ("bad_%d.h5" % off))
        open(q, "wb").write(bad)
        with pytest.raises(h5spec.SpecError):
            h5spec.check(q)
    q = str(tmp_path / "short.h5")
    open(q, "wb").write(raw[:-8])                                  # EOF address no longer the file size
    with pytest.raises(h5spec.SpecError):
        h5spec.check(q)


def _actdist_like(n, rng):
    return {"row": rng.integers(0, 30000, n).astype(np.int32), "col": rng.integers(0, 30000, n).astype(np.int32),
            "dist": rng.uniform(0, 5000, n).astype(np.float32), "prob": rng.uniform(0, 1, n).astype(np.float32)}


@pytest.mark.parametrize("n", [0, 1, 7, 3008, 100003])
def test_writer_output_passes_the_spec_checker(tmp_path, n):
    rng = np.random.default_rng(n)
    cols = _actdist_like(n, rng)
    p = str(tmp_path / "actdist.hdf5")
    hdf5.write_h5(p, cols)
    objs = h5spec.check(p)
    assert set(objs) == {"/", "/row", "/col", "/dist", "/prob"}
    raw = open(p, "rb").read()
    for k, a in cols.items():
        o = objs["/" + k]
        assert o["shape"] == (n,) and o["layout_class"] == 1 and o["elsize"] == 4
        addr, size = o["data"]
        assert size == a.nbytes and (size == 0 or raw[addr:addr + size] == a.tobytes())
        assert o["dtype_class"] == (0 if a.dtype.kind == "i" else 1)


def test_writer_groups_and_attributes_pass_the_spec_checker(tmp_path):
    p = str(tmp_path / "pop.hss")
    many = {"g/d%03d" % k: np.arange(k + 1, dtype=np.int64) for k in range(40)}      # several symbol nodes
    hdf5.write_h5(p, dict({"coordinates": np.zeros((5, 3, 3), np.float32), "radii": np.ones(5, np.float32),
                           "index/chrom": np.arange(5, dtype=np.int32),
                           "index/chromstr": np.array([b"chr1"] * 5, dtype="S10")}, **many),
                  attrs={"nbead": np.int64(5), "nstruct": np.int64(3), "version": np.int32(2)})
    objs = h5spec.check(p)
    assert objs["/g"]["kind"] == "group" and len([k for k in objs if k.startswith("/g/")]) == 40
    assert objs["/coordinates"]["shape"] == (5, 3, 3)
    assert sum(1 for m in objs["/"]["messages"] if m[0] == 0x0C) == 3


@needs_demo
def test_header_messages_equal_those_libhdf5_writes(tmp_path):
    """Same shape, same type -> the same dataspace / datatype / fill-value message bytes and the
    same layout message version and class as in the files the real library wrote."""
    real = h5spec.check(H.DEMO_HSS)
    p = str(tmp_path / "mine.h5")
    hdf5.write_h5(p, {"radii": np.zeros(3008, np.float32), "chrom": np.zeros(3008, np.int32),
                      "chromstr": np.array([b"chr1"] * 3008, dtype="S10"), "coordinates": np.zeros((3008, 100, 3), np.float32)})
    mine = h5spec.check(p)

    def msg(objs, path, mtype):
        return [m for m in objs[path]["messages"] if m[0] == mtype][0]
    for a, b in (("/radii", "/radii"), ("/chrom", "/index/chrom"), ("/chromstr", "/index/chromstr"),
                 ("/coordinates", "/coordinates")):
        for mtype in (0x01, 0x03):
            assert msg(mine, a, mtype)[1:] == msg(real, b, mtype)[1:], (a, hex(mtype))      # flags and bytes
    # the one contiguous array dataset of the real file: fill-value and layout messages too
    assert msg(mine, "/chromstr", 0x05)[1:] == msg(real, "/index/chromstr", 0x05)[1:]
    lm, lr = msg(mine, "/chromstr", 0x08)[2], msg(real, "/index/chromstr", 0x08)[2]
    assert lm[:2] == lr[:2] == bytes([3, 1])                       # version 3, contiguous
    assert struct.unpack_from("<Q", lm, 10) == struct.unpack_from("<Q", lr, 10) == (30080,)


@needs_demo
def test_reader_and_checker_agree_on_the_real_files():
    """The product's reader and the independent checker see the same objects in the files
    written by the real library."""
    for path in (H.DEMO_HSS, H.DEMO_HCS):
        objs = h5spec.check(path)
        with hdf5.open_h5(path) as f:
            for name, o in objs.items():
                if o["kind"] != "dataset" or o["dtype_class"] == 9:
                    continue
                node = f
                for part in name.strip("/").split("/"):
                    node = node[part]
                assert tuple(node.shape) == o["shape"], name
