"""Test infrastructure: an independent, read-only parser of the HDF5 file format subset that libhdf5 writes by
default for the reference's products (HDF5 File Format Specification v1.1 / 2.0: version-0 superblock, version-1
object headers, symbol-table groups = v1 B-tree + local heap + SNOD nodes, contiguous / compact layouts,
fixed-point / floating-point / string / compound / array datatypes).  libhdf5 and h5py are not in this image;
this reader is pinned against a file written by the real library (scipy's MATLAB v7.3 test file, see
tests/test_h5lite.py) and is then used to check the files the C++ writer (mara3_b200/csrc/h5lite.cpp) produces.

    f = H5File(path);  f.keys("/");  f.read("/solution/time");  f.dtype("/time_series")
"""
import struct
import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


class H5Error(ValueError):
    pass


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        self.base = None
        off = 0
        while off < len(self.buf):              # the superblock sits at 0, 512, 1024, ... (user block)
            if self.buf[off:off + 8] == SIGNATURE:
                self.base = off
                break
            off = 512 if off == 0 else off * 2
        if self.base is None:
            raise H5Error("no HDF5 signature")
        sb = self.buf[self.base:]
        self.superblock_version = sb[8]
        if self.superblock_version != 0:
            raise H5Error(f"superblock version {sb[8]} not supported")
        if sb[13] != 8 or sb[14] != 8:
            raise H5Error("only 8-byte offsets / lengths")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", sb, 16)
        base_addr, _, self.eof, _ = struct.unpack_from("<4Q", sb, 24)
        self.base = self.base if base_addr == 0 else base_addr
        name_off, self.root_header, cache, _ = struct.unpack_from("<QQII", sb, 56)
        self.root_cache_type = cache
        self.objects = {}

    # ---- low level -----------------------------------------------------------------------------------------
    def at(self, addr, n):
        a = self.base + addr
        if a + n > len(self.buf):
            raise H5Error(f"read past the end of the file at {addr}")
        return self.buf[a:a + n]

    def messages(self, addr):
        """[(type, flags, bytes)] of a version-1 object header, continuation blocks followed."""
        h = self.at(addr, 16)
        version, _, nmsg, refcount, size = struct.unpack_from("<BBHII", h, 0)
        if version != 1:
            raise H5Error(f"object header version {version} at {addr}")
        out, blocks = [], [(addr + 16, size)]
        while blocks and len(out) < nmsg:
            start, length = blocks.pop(0)
            p = 0
            while p + 8 <= length and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", self.at(start + p, 8), 0)
                data = self.at(start + p + 8, msize)
                p += 8 + msize
                if mtype == 0x0010:
                    blocks.append(struct.unpack_from("<QQ", data, 0))
                out.append((mtype, flags, data))
        return out

    def heap_name(self, heap_addr, offset):
        h = self.at(heap_addr, 32)
        if h[:4] != b"HEAP":
            raise H5Error("bad local heap signature")
        size, free, data = struct.unpack_from("<QQQ", h, 8)
        seg = self.at(data, size)
        end = seg.index(b"\0", offset)
        return seg[offset:end].decode()

    def group_entries(self, btree, heap):
        """{name: (object header address, cache type, scratch)} walking the v1 B-tree in key order."""
        out = {}

        def walk(addr):
            node = self.at(addr, 24)
            if node[:4] == b"SNOD":
                version, _, count = struct.unpack_from("<BBH", node, 4)
                for k in range(count):
                    e = self.at(addr + 8 + 40 * k, 40)
                    name_off, header, cache, _ = struct.unpack_from("<QQII", e, 0)
                    out[self.heap_name(heap, name_off)] = (header, cache, e[24:40])
                return
            if node[:4] != b"TREE":
                raise H5Error(f"bad B-tree node signature at {addr}")
            ntype, level, used = struct.unpack_from("<BBH", node, 4)
            if ntype != 0:
                raise H5Error("not a group B-tree")
            body = self.at(addr + 24, (2 * used + 1) * 8)
            keys = [struct.unpack_from("<Q", body, 16 * k)[0] for k in range(used + 1)]
            children = [struct.unpack_from("<Q", body, 16 * k + 8)[0] for k in range(used)]
            names = [self.heap_name(heap, k) for k in keys]
            assert names == sorted(names), "B-tree keys out of order"
            for c in children:
                walk(c)

        walk(btree)
        assert list(out) == sorted(out), "symbol table entries out of order"
        return out

    # ---- objects -------------------------------------------------------------------------------------------
    def lookup(self, path):
        addr = self.root_header
        for part in [p for p in path.split("/") if p]:
            entries = self.children(addr)
            if part not in entries:
                raise KeyError(path)
            addr = entries[part][0]
        return addr

    def children(self, addr):
        for mtype, flags, data in self.messages(addr):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", data, 0)
                return self.group_entries(btree, heap)
        raise H5Error("not a group")

    def is_group(self, path):
        return any(m[0] == 0x0011 for m in self.messages(self.lookup(path)))

    def keys(self, path="/"):
        return list(self.children(self.lookup(path)))

    def _dataset(self, path):
        shape = dtype = layout = None
        for mtype, flags, data in self.messages(self.lookup(path)):
            if mtype == 0x0001:
                version, rank, fl = data[0], data[1], data[2]
                if version == 1:
                    shape = struct.unpack_from(f"<{rank}Q", data, 8)
                elif version == 2:
                    shape = struct.unpack_from(f"<{rank}Q", data, 4)
                else:
                    raise H5Error("dataspace version")
            elif mtype == 0x0003:
                dtype, _ = parse_datatype(data, 0)
            elif mtype == 0x0008:
                version, cls = data[0], data[1]
                if version in (1, 2):               # libhdf5 <= 1.6: version, rank + 1, class, 5 reserved, address, sizes
                    if data[2] != 1:
                        raise H5Error("old-style layout that is not contiguous")
                    layout = ("contiguous", struct.unpack_from("<Q", data, 8)[0], None)
                    continue
                if version != 3:
                    raise H5Error(f"layout version {version}")
                if cls == 1:
                    layout = ("contiguous",) + struct.unpack_from("<QQ", data, 2)
                elif cls == 0:
                    n = struct.unpack_from("<H", data, 2)[0]
                    layout = ("compact", data[4:4 + n])
                else:
                    raise H5Error("chunked datasets are not supported by this reader")
        if shape is None or dtype is None or layout is None:
            raise H5Error(f"{path} is not a dataset")
        return tuple(shape), dtype, layout

    def shape(self, path):
        return self._dataset(path)[0]

    def dtype(self, path):
        return self._dataset(path)[1]

    def layout(self, path):
        return self._dataset(path)[2]

    def read(self, path):
        shape, dtype, layout = self._dataset(path)
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        nbytes = count * dtype.itemsize
        if layout[0] == "contiguous":
            raw = b"" if nbytes == 0 else self.at(layout[1], nbytes)
            if nbytes and layout[2] is not None:
                assert layout[2] == nbytes, "layout size does not match the dataspace x datatype"
        else:
            raw = layout[1][:nbytes]
        # numpy folds a sub-array dtype (HDF5 array type) into trailing dimensions
        a = np.frombuffer(raw, dtype=dtype, count=count).reshape(tuple(shape) + dtype.shape)
        return a.copy() if a.shape else a[()]


def parse_datatype(data, p):
    """numpy dtype described by the datatype message at data[p:], and the position after it."""
    cls_ver, b0, b1, b2, size = struct.unpack_from("<BBBBI", data, p)
    cls, version = cls_ver & 0x0F, cls_ver >> 4
    p += 8
    if cls == 0:                                    # fixed point
        if b0 & 1:
            raise H5Error("big-endian integers")
        signed = bool(b0 & 0x08)
        return np.dtype(("<i" if signed else "<u") + str(size)), p + 4
    if cls == 1:                                    # floating point
        if b0 & 1:
            raise H5Error("big-endian floats")
        return np.dtype("<f" + str(size)), p + 12
    if cls == 3:                                    # string
        return np.dtype("S" + str(size)), p
    if cls == 6:                                    # compound
        nmembers = b0 | (b1 << 8)
        names, formats, offsets = [], [], []
        for _ in range(nmembers):
            end = data.index(b"\0", p)
            name = data[p:end].decode()
            if version < 3:
                p += (end - p + 8) // 8 * 8         # name + terminator, padded to a multiple of 8
            else:
                p = end + 1
            if version == 1:
                offset = struct.unpack_from("<I", data, p)[0]
                p += 4 + 1 + 3 + 4 + 4 + 16
            elif version == 2:
                offset = struct.unpack_from("<I", data, p)[0]
                p += 4
            else:
                nb = 1 if size < 256 else (2 if size < 65536 else 4)
                offset = int.from_bytes(data[p:p + nb], "little")
                p += nb
            member, p = parse_datatype(data, p)
            names.append(name); formats.append(member); offsets.append(offset)
        return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), p
    if cls == 10:                                   # array
        rank = data[p]
        if version == 2:
            dims = struct.unpack_from(f"<{rank}I", data, p + 4)
            p += 4 + 8 * rank                       # sizes + permutation indices
        elif version == 3:
            dims = struct.unpack_from(f"<{rank}I", data, p + 1)
            p += 1 + 4 * rank
        else:
            raise H5Error("array datatype version")
        base, p = parse_datatype(data, p)
        return np.dtype((base, tuple(dims))), p
    raise H5Error(f"datatype class {cls} not supported")
