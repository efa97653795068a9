"""TEST INFRASTRUCTURE - an independent structural checker of classic-format HDF5 files.

Written from the HDF5 File Format Specification (version 1.1: superblock version 0,
symbol-table groups, local heaps, version-1 B-trees, version-1 object headers and the header
messages a plain ``h5py.create_dataset(name, data=array)`` produces), deliberately sharing NO
code with igm_b200/hdf5.py: it walks a file from the superblock, checks every invariant the
specification states for the structures it meets, and returns the tree of objects with their
raw header messages.  tests/test_hdf5_spec_cpu.py pins the checker itself on the reference's
demo files - written by the real libhdf5 through h5py - and then holds the files of
igm_b200.hdf5.write_h5 (what ``ActivationDistanceDB`` / the next A-step read with h5py,
igm/restraints/intra_hic.py:64-113, igm/steps/ActivationDistanceStep.py:145-151) to the same
rules, down to byte-identical dataspace / datatype messages.
"""
import struct

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
MSG_NAMES = {0x00: "nil", 0x01: "dataspace", 0x03: "datatype", 0x04: "fill_old", 0x05: "fill", 0x08: "layout",
             0x0B: "filters", 0x0C: "attribute", 0x10: "continuation", 0x11: "symbol_table", 0x12: "mtime"}


class SpecError(AssertionError):
    pass


def _req(cond, msg):
    if not cond:
        raise SpecError(msg)


class H5Spec:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        self.path = path
        self.objects = {}
        self._superblock()
        self._walk_group("/", self.root_header, self.root_btree, self.root_heap)

    # -- superblock (spec III.A, version 0) ---------------------------------------
    def _superblock(self):
        b = self.b
        _req(b[:8] == SIG, "format signature")
        ver, fsver, rgver, res0, shver, O, L, res1 = struct.unpack_from("<8B", b, 8)
        _req(ver == 0, "superblock version 0 expected, got %d" % ver)
        _req(fsver == 0 and rgver == 0 and shver == 0, "free-space / root-group / shared-header versions must be 0")
        _req(res0 == 0 and res1 == 0, "reserved superblock bytes must be zero")
        _req((O, L) == (8, 8), "8-byte offsets and lengths expected")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        _req(self.leaf_k > 0 and self.internal_k > 0, "group node K values must be positive")
        (flags,) = struct.unpack_from("<I", b, 20)
        _req(flags == 0, "file consistency flags must be 0 in a closed file")
        base, fs, eof, drv = struct.unpack_from("<QQQQ", b, 24)
        _req(base == 0, "base address 0 expected")
        _req(fs == UNDEF and drv == UNDEF, "no free-space manager / driver block expected")
        _req(eof == len(b), "end-of-file address %d != file size %d" % (eof, len(b)))
        self.eof = eof
        name_off, hdr, cache, resv = struct.unpack_from("<QQII", b, 56)
        _req(name_off == 0 and resv == 0, "root symbol-table entry: link name offset / reserved")
        _req(cache == 1, "root entry must cache the group's B-tree and heap (cache type 1)")
        self.root_header = hdr
        self.root_btree, self.root_heap = struct.unpack_from("<QQ", b, 80)

    def _in_file(self, addr, n, what):
        _req(addr != UNDEF and addr + n <= self.eof, "%s at %d (+%d) lies outside the file" % (what, addr, n))

    # -- object headers (spec IV.A.1.a, version 1) -----------------------------------
    def _messages(self, addr):
        b = self.b
        self._in_file(addr, 16, "object header")
        ver, resv, nmsg, refc, hsize = struct.unpack_from("<BBHII", b, addr)
        _req(ver == 1 and resv == 0, "version-1 object header expected")
        _req(refc >= 1, "object reference count")
        blocks = [(addr + 16, hsize)]
        self._in_file(addr + 16, hsize, "object header messages")
        out = []
        while blocks:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                _req(b[p + 5:p + 8] == b"\0\0\0", "reserved bytes of a message header")
                _req(msize % 8 == 0, "message data must be padded to 8 bytes (type 0x%x size %d)" % (mtype, msize))
                _req(p + 8 + msize <= end, "message 0x%x overruns its header block" % mtype)
                data = b[p + 8:p + 8 + msize]
                out.append((mtype, mflags, data))
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    self._in_file(caddr, clen, "continuation block")
                    blocks.append((caddr, clen))
                p += 8 + msize
            _req(p == end or end - p < 8, "gap inside an object header block")
        _req(len(out) == nmsg, "header says %d messages, found %d" % (nmsg, len(out)))
        for mtype, _, _ in out:
            _req(mtype in MSG_NAMES, "unknown header message type 0x%x" % mtype)
        return out

    # -- groups: symbol-table message, B-tree, symbol nodes, local heap (spec III.B-D) --
    def _heap(self, addr):
        b = self.b
        self._in_file(addr, 32, "local heap")
        _req(b[addr:addr + 4] == b"HEAP" and b[addr + 4] == 0 and b[addr + 5:addr + 8] == b"\0\0\0", "local heap header")
        size, free, data = struct.unpack_from("<QQQ", b, addr + 8)
        self._in_file(data, size, "local heap data segment")
        _req(size % 8 == 0, "heap data segment size must be a multiple of 8")
        # free list: (next, size) blocks inside the segment; 1 = end (also UNDEF in newer libraries)
        seen = 0
        while free not in (1, UNDEF):
            _req(free + 16 <= size and free % 8 == 0, "free block outside the heap segment")
            nxt, fsz = struct.unpack_from("<QQ", b, data + free)
            _req(fsz >= 16 and free + fsz <= size, "free block size")
            free = nxt
            seen += 1
            _req(seen < 1000, "free list loop")
        _req(b[data] == 0, "heap offset 0 must hold the empty string")
        return data, size

    def _name(self, heap, off):
        data, size = heap
        _req(off < size, "link name offset outside the heap")
        e = self.b.find(b"\0", data + off, data + size)
        _req(e >= 0, "unterminated link name")
        return self.b[data + off:e].decode("utf-8")

    def _walk_btree(self, addr, heap, out, depth=0):
        b = self.b
        self._in_file(addr, 24, "group B-tree node")
        _req(b[addr:addr + 4] == b"TREE", "B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        _req(ntype == 0, "group B-tree node type 0 expected")
        _req(used <= 2 * self.internal_k, "B-tree node holds more than 2K entries")
        full = 24 + 8 * (2 * self.internal_k + 1) + 8 * (2 * self.internal_k)
        self._in_file(addr, full, "B-tree node (allocated at full size)")
        left, right = struct.unpack_from("<QQ", b, addr + 8)
        if depth == 0:
            _req(left == UNDEF and right == UNDEF, "root B-tree node has no siblings")
        p = addr + 24
        keys = []
        for k in range(used):
            (key,) = struct.unpack_from("<Q", b, p)
            (child,) = struct.unpack_from("<Q", b, p + 8)
            keys.append(self._name(heap, key))
            if level > 0:
                self._walk_btree(child, heap, out, depth + 1)
            else:
                out.append(child)
            p += 16
        (last,) = struct.unpack_from("<Q", b, p)
        keys.append(self._name(heap, last))
        _req(keys == sorted(keys, key=lambda s: s.encode()), "B-tree keys must ascend")
        return keys

    def _walk_group(self, path, hdr, btree=None, heap_addr=None):
        msgs = self._messages(hdr)
        st = [m for m in msgs if m[0] == 0x11]
        _req(len(st) == 1, "%s: a group needs exactly one symbol-table message" % path)
        bt, hp = struct.unpack_from("<QQ", st[0][2], 0)
        if btree is not None:
            _req((bt, hp) == (btree, heap_addr), "cached B-tree / heap addresses differ from the symbol-table message")
        self.objects[path] = {"kind": "group", "messages": msgs}
        heap = self._heap(hp)
        snods = []
        self._walk_btree(bt, heap, snods)
        names = []
        for sn in snods:
            b = self.b
            self._in_file(sn, 8 + 40 * 2 * self.leaf_k, "symbol node (allocated at full size)")
            _req(b[sn:sn + 4] == b"SNOD" and b[sn + 4] == 1 and b[sn + 5] == 0, "symbol node header")
            (nsym,) = struct.unpack_from("<H", b, sn + 6)
            _req(nsym <= 2 * self.leaf_k, "symbol node holds more than 2K entries")
            for k in range(nsym):
                off, ohdr, cache, resv = struct.unpack_from("<QQII", b, sn + 8 + 40 * k)
                _req(resv == 0 and cache in (0, 1, 2), "symbol-table entry")
                nm = self._name(heap, off)
                _req(nm and "/" not in nm, "link name")
                names.append(nm)
                child = (path.rstrip("/") + "/" + nm)
                cm = self._messages(ohdr)
                if any(m[0] == 0x11 for m in cm):
                    if cache == 1:
                        cbt, chp = struct.unpack_from("<QQ", b, sn + 8 + 40 * k + 24)
                        self._walk_group(child, ohdr, cbt, chp)
                    else:
                        self._walk_group(child, ohdr)
                else:
                    self._dataset(child, cm)
        _req(names == sorted(names, key=lambda s: s.encode()), "%s: links must be stored in name order" % path)

    # -- datasets ---------------------------------------------------------------------
    def _dataset(self, path, msgs):
        by = {}
        for mtype, mflags, data in msgs:
            by.setdefault(mtype, []).append((mflags, data))
        for need in (0x01, 0x03, 0x08):
            _req(len(by.get(need, [])) == 1, "%s: exactly one %s message required" % (path, MSG_NAMES[need]))
        ds = by[0x01][0][1]
        ver, rank, flags = ds[0], ds[1], ds[2]
        _req(ver in (1, 2), "dataspace version")
        off = 8 if ver == 1 else 4
        dims = struct.unpack_from("<%dQ" % rank, ds, off)
        _req(flags & ~1 == 0 or ver == 2, "dataspace flags")
        dt = by[0x03][0][1]
        cls, dver = dt[0] & 0x0F, dt[0] >> 4
        _req(dver in (1, 2, 3), "datatype version")
        (esize,) = struct.unpack_from("<I", dt, 4)
        _req(esize > 0, "datatype size")
        if cls == 0:      # fixed point: bit offset, precision
            boff, prec = struct.unpack_from("<HH", dt, 8)
            _req(boff == 0 and prec == 8 * esize, "integer precision must fill the element")
            _req(dt[1] & 1 == 0, "little-endian expected")
        elif cls == 1:    # floating point
            boff, prec, eloc, esz, mloc, msz, bias = struct.unpack_from("<HHBBBBI", dt, 8)
            _req(boff == 0 and prec == 8 * esize, "float precision must fill the element")
            _req((eloc, esz, mloc, msz, bias) in ((23, 8, 0, 23, 127), (52, 11, 0, 52, 1023)), "IEEE float layout")
            _req(dt[1] & 1 == 0, "little-endian expected")
            _req(dt[2] == prec - 1, "sign bit location")
        lay = by[0x08][0][1]
        _req(lay[0] == 3, "data layout message version 3 expected")
        lclass = lay[1]
        n = 1
        for d in dims:
            n *= d
        info = {"kind": "dataset", "messages": msgs, "shape": tuple(dims), "elsize": esize, "dtype_class": cls,
                "dataspace_raw": bytes(ds), "datatype_raw": bytes(dt), "layout_class": lclass}
        if lclass == 1:
            addr, size = struct.unpack_from("<QQ", lay, 2)
            _req(size == n * esize, "%s: contiguous size %d != %d elements x %d bytes" % (path, size, n, esize))
            if size:
                self._in_file(addr, size, "%s: raw data" % path)      # (raw data need not be aligned)
            info["data"] = (addr, size)
            _req(0x0B not in by, "%s: a contiguous dataset cannot be filtered" % path)
        elif lclass == 2:
            ndim = lay[2]
            _req(ndim == rank + 1, "chunk dimensionality = rank + 1")
            cd = struct.unpack_from("<%dI" % ndim, lay, 11)
            _req(cd[-1] == esize, "last chunk dimension is the element size")
        elif lclass == 0:
            (size,) = struct.unpack_from("<H", lay, 2)
            _req(size == n * esize, "compact size")
        else:
            raise SpecError("unknown layout class %d" % lclass)
        for mflags, data in by.get(0x05, []):
            _req(data[0] in (1, 2, 3), "fill value message version")
        for mflags, data in by.get(0x0C, []):
            _req(data[0] in (1, 2, 3), "attribute message version")
        self.objects[path] = info


def check(path):
    """Parse and validate; returns {object path: info}."""
    return H5Spec(path).objects
