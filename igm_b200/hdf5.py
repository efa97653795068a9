"""Minimal classic-HDF5 reader and writer (no h5py / libhdf5 in this image).

Host-side I/O only: this is the on-disk boundary either side of the A-step hot
path, not part of the hot path itself.

Covers exactly what the IGM files at that boundary use:

* reader: superblock v0/v1, version-1 object headers (with continuation
  blocks), symbol-table groups (``TREE``/``SNOD``/``HEAP``), datasets with
  compact / contiguous / chunked (v1 B-tree) layout, deflate + shuffle
  filters, fixed-point / float / fixed-string / variable-length-string
  datatypes (global heap ``GCOL``), attributes v1-v3.  That is the format of
  ``demo/WTC11_HiC_2Mb.hcs``, ``demo/demo_sample_outputs/igm-model.hss.T`` and
  of ``actdist.hdf5`` as h5py writes it with default settings
  (reference: igm/steps/ActivationDistanceStep.py:285-289).
* writer: root-level groups one level deep, contiguous datasets of int/float
  types, scalar / 1-D numeric and fixed-string attributes.  Enough for
  ``actdist.hdf5`` (``row,col: int32``, ``dist,prob: float32``; read back by
  igm/restraints/intra_hic.py:64-113 and ActivationDistanceStep.py:145-151)
  and for synthetic ``.hss`` / ``.hcs`` test files.

If ``h5py`` is importable, ``open_h5``/``write_h5`` prefer it.
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5FormatError(IOError):
    pass


# --------------------------------------------------------------------------
# reader
# --------------------------------------------------------------------------

class _Datatype:
    __slots__ = ("cls", "size", "np_dtype", "vlen_string", "base")

    def __init__(self):
        self.cls = -1
        self.size = 0
        self.np_dtype = None
        self.vlen_string = False
        self.base = None


def _parse_datatype(buf: bytes, off: int = 0) -> _Datatype:
    b0, b1, b2, b3, size = struct.unpack_from("<BBBBI", buf, off)
    cls = b0 & 0x0F
    dt = _Datatype()
    dt.cls = cls
    dt.size = size
    if cls == 0:  # fixed point
        order = ">" if (b1 & 1) else "<"
        signed = bool(b1 & 0x08)
        dt.np_dtype = np.dtype("%s%s%d" % (order, "i" if signed else "u", size))
    elif cls == 1:  # floating point
        order = ">" if (b1 & 1) else "<"
        dt.np_dtype = np.dtype("%sf%d" % (order, size))
    elif cls == 3:  # fixed-length string
        dt.np_dtype = np.dtype("S%d" % size)
    elif cls == 9:  # variable length
        vtype = b1 & 0x0F
        dt.vlen_string = (vtype == 1)
        dt.base = _parse_datatype(buf, off + 8)
        dt.np_dtype = np.dtype("O")
    elif cls == 8:  # enum (h5py bools): base type follows
        dt.base = _parse_datatype(buf, off + 8)
        dt.np_dtype = dt.base.np_dtype
    else:
        raise Hdf5FormatError("unsupported HDF5 datatype class %d" % cls)
    return dt


class Dataset:
    """Lazy handle on one dataset; ``ds[()]`` / ``ds[:]`` / ``ds[a:b]`` read it."""

    def __init__(self, f: "File", name: str, msgs):
        self._f = f
        self.name = name
        self.attrs: Dict[str, Any] = {}
        self.shape: Tuple[int, ...] = ()
        self._dt: Optional[_Datatype] = None
        self._layout = None
        self._filters: List[Tuple[int, Tuple[int, ...]]] = []
        for mtype, data in msgs:
            if mtype == 0x01:
                self.shape = _parse_dataspace(data, f._L)
            elif mtype == 0x03:
                self._dt = _parse_datatype(data)
            elif mtype == 0x08:
                self._layout = data
            elif mtype == 0x0B:
                self._filters = _parse_filters(data)
            elif mtype == 0x0C:
                k, v = f._parse_attribute(data)
                self.attrs[k] = v
        if self._dt is None or self._layout is None:
            raise Hdf5FormatError("object %s is not a dataset" % name)

    @property
    def dtype(self):
        return self._dt.np_dtype

    def __len__(self):
        return self.shape[0] if self.shape else 0

    def read(self) -> np.ndarray:
        f = self._f
        lay = self._layout
        ver = lay[0]
        n = int(np.prod(self.shape)) if self.shape else 1
        esize = self._dt.size
        if ver == 3:
            cls = lay[1]
            if cls == 0:  # compact
                (sz,) = struct.unpack_from("<H", lay, 2)
                raw = lay[4:4 + sz]
            elif cls == 1:  # contiguous
                addr, sz = struct.unpack_from("<QQ", lay, 2)
                raw = b"" if addr == _UNDEF else f._read(addr, n * esize)
                if addr == _UNDEF:
                    raw = b"\0" * (n * esize)
            elif cls == 2:  # chunked
                ndim = lay[2]
                (btree,) = struct.unpack_from("<Q", lay, 3)
                cdims = struct.unpack_from("<%dI" % ndim, lay, 11)
                return self._read_chunked(btree, cdims[:-1])
            else:
                raise Hdf5FormatError("bad layout class %d" % cls)
        elif ver in (1, 2):
            ndim, cls = lay[1], lay[2]
            off = 8
            addr = _UNDEF
            if cls != 0:
                (addr,) = struct.unpack_from("<Q", lay, off)
                off += 8
            dims = struct.unpack_from("<%dI" % ndim, lay, off)
            off += 4 * ndim
            if cls == 0:
                (sz,) = struct.unpack_from("<I", lay, off)
                raw = lay[off + 4:off + 4 + sz]
            elif cls == 1:
                raw = f._read(addr, n * esize)
            else:
                return self._read_chunked(addr, dims[:-1])
        else:
            raise Hdf5FormatError("unsupported data layout version %d" % ver)
        return self._decode(raw, self.shape)

    def _decode(self, raw: bytes, shape) -> np.ndarray:
        dt = self._dt
        n = int(np.prod(shape)) if shape else 1
        if dt.cls == 9:
            out = np.empty(n, dtype=object)
            for k in range(n):
                ln, addr, idx = struct.unpack_from("<IQI", raw, 16 * k)
                obj = self._f._global_heap_object(addr, idx)[:ln] if addr not in (0, _UNDEF) else b""
                out[k] = obj.decode("utf-8") if dt.vlen_string else obj
            return out.reshape(shape)
        arr = np.frombuffer(raw, dtype=dt.np_dtype, count=n).reshape(shape)
        if dt.np_dtype.kind in "iuf" and dt.np_dtype.byteorder == ">":
            arr = arr.astype(dt.np_dtype.newbyteorder("<"))
        return arr.copy()

    def _chunk_layout(self):
        """(btree address, chunk dims) of a chunked dataset, else None."""
        lay = self._layout
        if lay[0] == 3 and lay[1] == 2:
            ndim = lay[2]
            (btree,) = struct.unpack_from("<Q", lay, 3)
            cdims = struct.unpack_from("<%dI" % ndim, lay, 11)
            return btree, cdims[:-1]
        if lay[0] in (1, 2) and lay[2] == 2:
            ndim = lay[1]
            (addr,) = struct.unpack_from("<Q", lay, 8)
            dims = struct.unpack_from("<%dI" % ndim, lay, 16)
            return addr, dims[:-1]
        return None

    def raw_extents(self):
        """[(file_offset, nbytes, dest_byte_offset)] when the stored bytes ARE the array
        (little-endian numeric type, no filter, and either a contiguous layout or chunks
        that span all trailing dimensions, i.e. whole leading-axis slabs - what the
        production .hss holds: pack_beads x nstruct x 3, igm/steps/ModelingStep.py:753-760);
        None otherwise.  Lets a loader pread() the coordinates straight into pinned memory."""
        dt = self._dt
        if dt.cls == 9 or dt.np_dtype.kind not in "iuf" or dt.np_dtype.byteorder == ">" or self._filters:
            return None
        if not self.shape:
            return None
        esize = dt.size
        rowb = esize * int(np.prod(self.shape[1:])) if len(self.shape) > 1 else esize
        base = self._f._base
        cl = self._chunk_layout()
        if cl is None:
            lay = self._layout
            if lay[0] == 3 and lay[1] == 1:
                addr, _sz = struct.unpack_from("<QQ", lay, 2)
            elif lay[0] in (1, 2) and lay[2] == 1:
                (addr,) = struct.unpack_from("<Q", lay, 8)
            else:
                return None
            if addr == _UNDEF:
                return None
            return [(addr + base, rowb * self.shape[0], 0)]
        btree, cdims = cl
        if btree == _UNDEF or tuple(cdims[1:]) != tuple(self.shape[1:]):
            return None
        out = []
        for csize, fmask, offs, addr in self._f._iter_chunks(btree, len(self.shape)):
            if fmask or any(offs[1:]):
                return None
            rows = min(cdims[0], self.shape[0] - offs[0])
            if rows > 0:
                out.append((addr + base, rows * rowb, offs[0] * rowb))
        return out

    def iter_chunks(self):
        """Yields (offsets, ndarray) for every stored chunk of a chunked dataset, clipped
        to the dataset extent (one pseudo-chunk at offset 0 for other layouts).  Lets a
        loader stream a large .hss straight to the GPU without holding it in host memory
        (the production files are chunked pack_beads x nstruct x 3,
        igm/steps/ModelingStep.py:753-760)."""
        cl = self._chunk_layout()
        if cl is None or cl[0] == _UNDEF:
            yield tuple(0 for _ in self.shape), self.read()
            return
        btree, cdims = cl
        for offs, chunk in self._decoded_chunks(btree, cdims):
            sl = tuple(slice(0, min(cdims[d], self.shape[d] - offs[d])) for d in range(len(self.shape)))
            yield tuple(offs[:len(self.shape)]), chunk[sl]

    def _decoded_chunks(self, btree: int, cdims):
        f = self._f
        ndim = len(self.shape)
        esize = self._dt.size
        for csize, fmask, offs, addr in f._iter_chunks(btree, ndim):
            raw = f._read(addr, csize)
            for pos in range(len(self._filters) - 1, -1, -1):
                fid, cd = self._filters[pos]
                if fmask & (1 << pos):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    k = cd[0] if cd else esize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    m = len(a) // k
                    raw = a[:m * k].reshape(k, m).T.tobytes() + a[m * k:].tobytes()
                elif fid == 3:
                    raw = raw[:-4]  # fletcher32 checksum trailer
                else:
                    raise Hdf5FormatError("unsupported HDF5 filter id %d" % fid)
            yield offs, np.frombuffer(raw, dtype=self._dt.np_dtype,
                                      count=int(np.prod(cdims))).reshape(cdims)

    def _read_chunked(self, btree: int, cdims) -> np.ndarray:
        f = self._f
        shape = self.shape
        ndim = len(shape)
        if self._dt.cls == 9:
            raise Hdf5FormatError("chunked vlen datasets unsupported")
        out = np.zeros(shape, dtype=self._dt.np_dtype)
        if btree == _UNDEF:
            return out
        esize = self._dt.size
        for csize, fmask, offs, addr in f._iter_chunks(btree, ndim):
            raw = f._read(addr, csize)
            for pos in range(len(self._filters) - 1, -1, -1):
                fid, cd = self._filters[pos]
                if fmask & (1 << pos):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    k = cd[0] if cd else esize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    m = len(a) // k
                    raw = a[:m * k].reshape(k, m).T.tobytes() + a[m * k:].tobytes()
                elif fid == 3:
                    raw = raw[:-4]  # fletcher32 checksum trailer
                else:
                    raise Hdf5FormatError("unsupported HDF5 filter id %d" % fid)
            chunk = np.frombuffer(raw, dtype=self._dt.np_dtype,
                                  count=int(np.prod(cdims))).reshape(cdims)
            sl_out, sl_in = [], []
            for d in range(ndim):
                lo = offs[d]
                hi = min(lo + cdims[d], shape[d])
                sl_out.append(slice(lo, hi))
                sl_in.append(slice(0, hi - lo))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]
        return out

    def __getitem__(self, key):
        arr = self.read()
        if key is Ellipsis or (isinstance(key, tuple) and key == ()):
            if arr.shape == ():
                return arr[()]
            return arr
        return arr[key]


def _parse_dataspace(data: bytes, L: int) -> Tuple[int, ...]:
    ver, rank, flags = data[0], data[1], data[2]
    if ver == 1:
        off = 8
    elif ver == 2:
        off = 4
        if data[3] == 2:  # null dataspace
            return (0,)
    else:
        raise Hdf5FormatError("unsupported dataspace version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, data, off)) if rank else ()


def _parse_filters(data: bytes):
    ver, nf = data[0], data[1]
    out = []
    off = 8 if ver == 1 else 2
    for _ in range(nf):
        if ver == 1:
            fid, nlen, flags, ncd = struct.unpack_from("<HHHH", data, off)
            off += 8
            off += (nlen + 7) // 8 * 8
        else:
            (fid,) = struct.unpack_from("<H", data, off)
            off += 2
            nlen = 0
            if fid >= 256:
                (nlen,) = struct.unpack_from("<H", data, off)
                off += 2
            flags, ncd = struct.unpack_from("<HH", data, off)
            off += 4 + nlen
        cd = struct.unpack_from("<%dI" % ncd, data, off)
        off += 4 * ncd
        if ver == 1 and (ncd & 1):
            off += 4
        out.append((fid, cd))
    return out


class Group:
    def __init__(self, f: "File", name: str, msgs):
        self._f = f
        self.name = name
        self.attrs: Dict[str, Any] = {}
        self._links: Dict[str, int] = {}
        for mtype, data in msgs:
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", data, 0)
                self._links = f._read_symbol_table(btree, heap)
            elif mtype == 0x0C:
                k, v = f._parse_attribute(data)
                self.attrs[k] = v

    def keys(self):
        return list(self._links.keys())

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self._links)

    def __getitem__(self, path: str):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(path)
            node = node._f._open_object(node._links[part], node.name.rstrip("/") + "/" + part)
        return node


class File(Group):
    """Read-only view of a classic-format HDF5 file (memory-mapped: opening a 3.6 GB .hss
    touches only the metadata pages)."""

    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise ValueError("igm_b200.hdf5.File is read-only; use write_h5() to write")
        import mmap
        with open(path, "rb") as fh:
            size = os.fstat(fh.fileno()).st_size
            self._buf = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) if size else b""
        self.filename = path
        b = self._buf
        if b[:8] != _SIG:
            raise Hdf5FormatError("%s: not an HDF5 file" % path)
        ver = b[8]
        if ver not in (0, 1):
            raise Hdf5FormatError("%s: superblock version %d unsupported (classic only)" % (path, ver))
        self._O, self._L = b[13], b[14]
        if (self._O, self._L) != (8, 8):
            raise Hdf5FormatError("only 8-byte offsets/lengths supported")
        off = 24 if ver == 0 else 28
        self._base, _fs, _eof, _drv = struct.unpack_from("<QQQQ", b, off)
        off += 32
        _lno, root_hdr, cache, _r = struct.unpack_from("<QQII", b, off)
        self._obj_cache: Dict[int, Any] = {}
        self._gheap_cache: Dict[int, Dict[int, bytes]] = {}
        Group.__init__(self, self, "/", self._read_object_header(root_hdr))

    # context manager / h5py-like surface
    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        b, self._buf = self._buf, b""
        try:
            if hasattr(b, "close"):
                b.close()
        except (BufferError, ValueError):       # a view is still alive somewhere: let the GC do it
            pass

    def _read(self, addr: int, n: int) -> bytes:
        a = addr + self._base
        if a + n > len(self._buf):
            raise Hdf5FormatError("read past end of file")
        return self._buf[a:a + n]

    def _read_object_header(self, addr: int):
        b = self._buf
        a = addr + self._base
        ver = b[a]
        if ver != 1:
            raise Hdf5FormatError("object header version %d unsupported (classic v1 only)" % ver)
        nmsg, _ref, hsize = struct.unpack_from("<HII", b, a + 2)
        blocks = [(a + 16, hsize)]
        msgs = []
        bi = 0
        while bi < len(blocks) and len(msgs) < nmsg:
            p, size = blocks[bi]
            end = p + size
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    coff, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((coff + self._base, clen))
                msgs.append((mtype, data))
            bi += 1
        return msgs

    def _open_object(self, addr: int, name: str):
        if addr in self._obj_cache:
            return self._obj_cache[addr]
        msgs = self._read_object_header(addr)
        types = {m for m, _ in msgs}
        obj = Group(self, name, msgs) if 0x11 in types else Dataset(self, name, msgs)
        self._obj_cache[addr] = obj
        return obj

    def _heap_data_addr(self, heap: int) -> int:
        b = self._buf
        a = heap + self._base
        if b[a:a + 4] != b"HEAP":
            raise Hdf5FormatError("bad local heap signature")
        _sz, _free, daddr = struct.unpack_from("<QQQ", b, a + 8)
        return daddr + self._base

    def _read_symbol_table(self, btree: int, heap: int) -> Dict[str, int]:
        b = self._buf
        hdata = self._heap_data_addr(heap)
        links: Dict[str, int] = {}

        def name_at(off):
            s = hdata + off
            e = b.find(b"\0", s)        # (mmap has find, not index)
            if e < 0:
                raise Hdf5FormatError("unterminated link name")
            return b[s:e].decode("utf-8")

        def walk(addr):
            a = addr + self._base
            if b[a:a + 4] != b"TREE":
                raise Hdf5FormatError("bad B-tree signature")
            ntype, level, used = struct.unpack_from("<BBH", b, a + 4)
            p = a + 8 + 16
            for k in range(used):
                (child,) = struct.unpack_from("<Q", b, p + 8)
                p += 16
                if level > 0:
                    walk(child)
                else:
                    s = child + self._base
                    if b[s:s + 4] != b"SNOD":
                        raise Hdf5FormatError("bad symbol node signature")
                    (nsym,) = struct.unpack_from("<H", b, s + 6)
                    q = s + 8
                    for _ in range(nsym):
                        lno, ohdr = struct.unpack_from("<QQ", b, q)
                        links[name_at(lno)] = ohdr
                        q += 40

        walk(btree)
        return links

    def _iter_chunks(self, btree: int, ndim: int):
        b = self._buf
        keysz = 8 + 8 * (ndim + 1)

        def walk(addr):
            a = addr + self._base
            if b[a:a + 4] != b"TREE":
                raise Hdf5FormatError("bad chunk B-tree signature")
            ntype, level, used = struct.unpack_from("<BBH", b, a + 4)
            p = a + 24
            for k in range(used):
                csize, fmask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from("<%dQ" % (ndim + 1), b, p + 8)
                (child,) = struct.unpack_from("<Q", b, p + keysz)
                p += keysz + 8
                if level > 0:
                    yield from walk(child)
                else:
                    yield csize, fmask, offs[:ndim], child

        yield from walk(btree)

    def _global_heap_object(self, addr: int, idx: int) -> bytes:
        col = self._gheap_cache.get(addr)
        if col is None:
            b = self._buf
            a = addr + self._base
            if b[a:a + 4] != b"GCOL":
                raise Hdf5FormatError("bad global heap signature")
            (csize,) = struct.unpack_from("<Q", b, a + 8)
            col = {}
            p = a + 16
            end = a + csize
            while p + 16 <= end:
                oidx, _rc, _r, osz = struct.unpack_from("<HHIQ", b, p)
                if oidx == 0:
                    break
                col[oidx] = b[p + 16:p + 16 + osz]
                p += 16 + (osz + 7) // 8 * 8
            self._gheap_cache[addr] = col
        return col[idx]

    def _parse_attribute(self, data: bytes):
        ver = data[0]
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from("<HHH", data, 2)
            p = 8
            pad = lambda x: (x + 7) // 8 * 8
        elif ver in (2, 3):
            nsz, tsz, ssz = struct.unpack_from("<HHH", data, 2)
            p = 8 if ver == 2 else 9
            pad = lambda x: x
        else:
            raise Hdf5FormatError("attribute version %d unsupported" % ver)
        name = data[p:p + nsz].split(b"\0")[0].decode("utf-8")
        p += pad(nsz)
        dt = _parse_datatype(data, p)
        p += pad(tsz)
        shape = _parse_dataspace(data[p:p + ssz], self._L)
        p += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        if dt.cls == 9:
            vals = []
            for k in range(n):
                ln, addr, idx = struct.unpack_from("<IQI", data, p + 16 * k)
                raw = self._global_heap_object(addr, idx)[:ln]
                vals.append(raw.decode("utf-8") if dt.vlen_string else raw)
            val = vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
        else:
            arr = np.frombuffer(data, dtype=dt.np_dtype, count=n, offset=p).reshape(shape)
            val = arr[()] if shape == () else arr.copy()
        return name, val


def _real_h5py():
    """h5py if genuinely installed (test harnesses may register stub modules)."""
    try:
        import h5py  # type: ignore
    except ImportError:
        return None
    return h5py if isinstance(getattr(h5py, "__version__", None), str) else None


def open_h5(path: str):
    """Open ``path`` read-only with h5py when present, else the built-in reader."""
    h5py = _real_h5py()
    if h5py is not None:
        return h5py.File(path, "r")
    return File(path, "r")


# --------------------------------------------------------------------------
# writer
# --------------------------------------------------------------------------

def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10 | 0, bits0, 0, 0, dt.itemsize) + \
            struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "f":
        if dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            sign = 31
        elif dt.itemsize == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            sign = 63
        else:
            raise ValueError("unsupported float size")
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, dt.itemsize) + props
    if dt.kind == "S":
        # null-padded ASCII fixed string
        return struct.pack("<BBBBI", 0x10 | 3, 0x01, 0, 0, dt.itemsize)
    raise ValueError("unsupported dtype for HDF5 writer: %r" % dt)


def _dataspace_message(shape, maxdims: bool = False) -> bytes:
    """Version-1 simple dataspace.  ``maxdims``: also store the maximum dimensions (= the
    current ones), as libhdf5 does for every dataset h5py creates (flags bit 0) - the demo
    files written by the real library show it (tests/test_hdf5_spec_cpu.py)."""
    shape = tuple(int(s) for s in shape)
    dims = b"".join(struct.pack("<Q", s) for s in shape)
    if maxdims and shape:
        return struct.pack("<BBBB4x", 1, len(shape), 1, 0) + dims + dims
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + dims


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _attribute_message(name: str, value) -> bytes:
    if isinstance(value, (bytes, str)):
        raw = value.encode("utf-8") if isinstance(value, str) else value
        arr = np.array(raw, dtype="S%d" % max(1, len(raw)))
    else:
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = arr.astype("S")
        if arr.dtype == np.bool_:
            arr = arr.astype(np.int8)
    nm = name.encode("utf-8") + b"\0"
    dtm = _dtype_message(arr.dtype)
    dsm = _dataspace_message(arr.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(dsm))
    body += _pad8(nm) + _pad8(dtm) + _pad8(dsm) + np.ascontiguousarray(arr).tobytes()
    return _message(0x0C, body)


def _object_header(messages: List[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _Writer:
    """Append-only image builder; every structure is 8-byte aligned."""

    GROUP_LEAF_K = 4      # symbol node holds up to 2K entries
    GROUP_INTERNAL_K = 16  # B-tree node holds up to 2K children

    def __init__(self):
        # the image is kept as a list of byte blocks and streamed to the file at the
        # end: large datasets are written straight from the arrays' own memory
        self.parts = []
        self.pos = 0

    def alloc(self, data) -> int:
        pad = -self.pos % 8
        if pad:
            self.parts.append(b"\0" * pad)
            self.pos += pad
        addr = self.pos
        self.parts.append(data)
        self.pos += len(data)
        return addr

    def write_dataset(self, arr: np.ndarray, attrs: Optional[dict] = None) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        if arr.dtype.kind == "U":
            arr = arr.astype("S")
        raw = memoryview(arr.reshape(-1).view(np.uint8)) if arr.size else b""
        daddr = self.alloc(raw) if len(raw) else _UNDEF
        msgs = [
            _message(0x01, _dataspace_message(arr.shape, maxdims=True)),
            _message(0x03, _dtype_message(arr.dtype), flags=1),
            # fill value, version 2: allocate late (2), write the fill value if set (2), defined
            # (1) with size 0 = the library default - byte for byte what libhdf5 stores for a
            # contiguous dataset created through h5py
            _message(0x05, struct.pack("<BBBBI", 2, 2, 2, 1, 0), flags=1),
            _message(0x08, struct.pack("<BBQQ", 3, 1, daddr, len(raw))),
        ]
        for k, v in (attrs or {}).items():
            msgs.append(_attribute_message(k, v))
        return self.alloc(_object_header(msgs))

    def write_group(self, links: Dict[str, int], attrs: Optional[dict] = None) -> Tuple[int, int, int]:
        """links: name -> object header address.  Returns (header, btree, heap)."""
        names = sorted(links.keys(), key=lambda s: s.encode("utf-8"))
        # local heap: offset 0 holds the empty string
        heap_data = bytearray(b"\0" * 8)
        offs = {}
        for nm in names:
            offs[nm] = len(heap_data)
            heap_data += nm.encode("utf-8") + b"\0"
            heap_data += b"\0" * (-len(heap_data) % 8)
        # a free block needs >= 16 bytes; append one so the heap is well formed
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)  # next free = 1 (none), size = 16
        hd_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, hd_addr))
        # symbol nodes, each up to 2K entries, allocated at full size
        cap = 2 * self.GROUP_LEAF_K
        snods = []
        for s in range(0, max(1, len(names)), cap):
            part = names[s:s + cap]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for nm in part:
                body += struct.pack("<QQII16x", offs[nm], links[nm], 0, 0)
            body += b"\0" * (40 * (cap - len(part)))
            snods.append((self.alloc(body), offs[part[-1]] if part else 0))
        if len(snods) > 2 * self.GROUP_INTERNAL_K:
            raise ValueError("too many links for the single-level group writer")
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), _UNDEF, _UNDEF)
        node += struct.pack("<Q", 0)  # key 0: empty string
        for addr, last_key in snods:
            node += struct.pack("<QQ", addr, last_key)
        full = 24 + 8 + 16 * (2 * self.GROUP_INTERNAL_K)
        node += b"\0" * (full - len(node))
        bt_addr = self.alloc(node)
        msgs = [_message(0x11, struct.pack("<QQ", bt_addr, heap_addr))]
        for k, v in (attrs or {}).items():
            msgs.append(_attribute_message(k, v))
        hdr = self.alloc(_object_header(msgs))
        return hdr, bt_addr, heap_addr


def write_h5(path: str, datasets: Dict[str, Any], attrs: Optional[dict] = None) -> None:
    """Write a classic HDF5 file.

    ``datasets`` maps ``"name"`` or ``"group/name"`` (one level) to arrays; a
    value may also be ``(array, attrs_dict)``.  ``attrs`` are root attributes.
    Output layout matches what h5py's ``create_dataset(name, data=arr)`` yields
    for these dtypes: contiguous, unfiltered, little-endian
    (reference writer: igm/steps/ActivationDistanceStep.py:285-289).
    """
    h5py = _real_h5py()
    if h5py is not None:
        with h5py.File(path, "w") as f:
            for k, v in (attrs or {}).items():
                f.attrs[k] = v
            for name, val in datasets.items():
                a, at = val if isinstance(val, tuple) else (val, None)
                d = f.create_dataset(name, data=np.asarray(a))
                for k, v in (at or {}).items():
                    d.attrs[k] = v
        return

    w = _Writer()
    w.alloc(b"\0" * 96)  # superblock v0 placeholder (56 + 40-byte root entry)
    root_links: Dict[str, int] = {}
    groups: Dict[str, Dict[str, int]] = {}
    for name, val in datasets.items():
        a, at = val if isinstance(val, tuple) else (val, None)
        addr = w.write_dataset(np.asarray(a), at)
        parts = [p for p in name.split("/") if p]
        if len(parts) == 1:
            root_links[parts[0]] = addr
        elif len(parts) == 2:
            groups.setdefault(parts[0], {})[parts[1]] = addr
        else:
            raise ValueError("write_h5 supports at most one group level: %s" % name)
    for g, links in groups.items():
        hdr, _, _ = w.write_group(links)
        root_links[g] = hdr
    root_hdr, bt, heap = w.write_group(root_links, attrs)
    eof = w.pos + (-w.pos % 8)
    if eof > w.pos:
        w.parts.append(b"\0" * (eof - w.pos))
    sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0,
                            _Writer.GROUP_LEAF_K, _Writer.GROUP_INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", bt, heap)
    assert len(sb) == 96
    w.parts[0] = sb
    with open(path, "wb") as fh:
        for part in w.parts:
            fh.write(part)
