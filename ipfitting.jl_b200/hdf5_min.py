"""Minimal HDF5 reader / writer for the one thing the reference stores in `<dbpath>_kron.h5`: a single contiguous
float64 matrix, dataset "A" in the root group (`_savemath5` / `_loadmath5`, `/root/reference/src/lsq_db.jl:111-124`).

Written from the HDF5 File Format Specification (version 0 superblock, version 1 object headers, symbol-table groups:
what libhdf5 emits with default library-version bounds, i.e. what `h5open(fname, "w")` of HDF5.jl produces).  No HDF5
library exists in the build image, so this module is checked against the specification and against itself only
(tests/test_hdf5_min.py) -- NOT against libhdf5; DESIGN.md §8 says so.  Not supported, with a clear error: chunked or
compressed datasets, files written with libver = latest (version 2 object headers), big-endian or non-IEEE types.

Layout of a written file (all addresses absolute, base address 0):
    0     superblock v0 (96 bytes) with the root symbol-table entry
    96    root object header v1: one symbol-table message (B-tree, local heap)
    136   B-tree node (type 0 = group, level 0, one child)
    680   local heap header, 712 its data segment ("" and the dataset name)
    728   symbol-table node with one entry
    1056  dataset object header v1: dataspace, datatype, contiguous layout
    2048  the data (C order over the HDF5 dims = Fortran order over the Julia dims)
"""
from __future__ import annotations

import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K, _INTERNAL_K = 4, 16
_DATA_ADDR = 2048
# IEEE binary64, little endian: class 1 (floating point) version 1; bit field 0x20 = mantissa normalisation "implied
# msb", sign bit 63; size 8; properties: bit offset 0, precision 64, exponent at 52 (11 bits), mantissa at 0 (52 bits),
# bias 1023
_F64LE = bytes([0x11, 0x20, 0x3F, 0x00]) + struct.pack("<I", 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


class HDF5FormatError(RuntimeError):
    pass


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    # version 1, reserved, number of messages, reference count, header size; 4 bytes of padding align the messages
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def write_matrix(fname: str, A: np.ndarray, name: str = "A") -> None:
    """Write the Julia matrix `A` (m x n) the way HDF5.jl does: HDF5 dims (n, m), row-major = A's column-major bytes."""
    A = np.asarray(A, dtype=np.float64)
    if A.ndim != 2:
        raise ValueError("a matrix is expected")
    nm = name.encode()
    if not nm or b"\0" in nm or b"/" in nm:
        raise ValueError("bad dataset name")
    m, n = A.shape
    rawT = np.asfortranarray(A).T             # C-contiguous view with the HDF5 dims: no copy of a matrix already in Julia order
    nbytes = rawT.nbytes
    hdf_dims = (n, m)

    addr_root, addr_btree = 96, 136
    btree_size = 24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8
    addr_heap = addr_btree + btree_size
    heap_data = _pad8(b"\0") + _pad8(nm + b"\0")
    addr_heap_data = addr_heap + 32
    addr_snod = addr_heap_data + len(heap_data)
    snod_size = 8 + 2 * _LEAF_K * 40
    addr_dset = addr_snod + snod_size
    name_off = 8

    dataspace = struct.pack("<BBB5x", 1, len(hdf_dims), 0) + b"".join(struct.pack("<Q", d) for d in hdf_dims)
    layout = struct.pack("<BBQQ", 3, 1, _DATA_ADDR, nbytes)          # version 3, class 1 = contiguous
    dset = _object_header([_message(0x0001, dataspace), _message(0x0003, _F64LE, flags=1), _message(0x0008, layout)])
    if addr_dset + len(dset) > _DATA_ADDR:
        raise ValueError("dataset name too long for the fixed layout")
    root = _object_header([_message(0x0011, struct.pack("<QQ", addr_btree, addr_heap))])
    assert addr_root + len(root) == addr_btree

    eof = _DATA_ADDR + nbytes
    sb = SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", _LEAF_K, _INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, addr_root, 1, 0) + struct.pack("<QQ", addr_btree, addr_heap)      # root entry, cached
    assert len(sb) == 96

    btree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, addr_snod, name_off)
    btree += b"\0" * (btree_size - len(btree))
    heap = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, addr_heap_data)   # free list: 1 = none (H5HL_FREE_NULL)
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, 1) + struct.pack("<QQII16x", name_off, addr_dset, 0, 0)
    snod += b"\0" * (snod_size - len(snod))

    head = sb + root + btree + heap + heap_data + snod + dset
    head += b"\0" * (_DATA_ADDR - len(head))
    with open(fname, "wb") as f:
        f.write(head)
        rawT.tofile(f)


# ---------------------------------------------------------------------------------------------------------------
class _File:
    def __init__(self, buf):
        self.b = buf                     # bytes or mmap: slices are bytes, struct reads in place
        if buf[:8] != SIGNATURE:
            raise HDF5FormatError("not an HDF5 file (signature at offset 0 expected)")
        ver = buf[8]
        if ver not in (0, 1):
            raise HDF5FormatError(f"superblock version {ver}: files written with libver = latest are not supported")
        if buf[13] != 8 or buf[14] != 8:
            raise HDF5FormatError("only 8-byte offsets and lengths are supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", buf, 16)
        p = 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", buf, p)
        if self.eof > len(buf) + self.base:
            raise HDF5FormatError("truncated file (end-of-file address beyond the file size)")
        p += 32
        _, self.root_header, cache, _ = struct.unpack_from("<QQII", buf, p)
        self.root_cache = struct.unpack_from("<QQ", buf, p + 24) if cache == 1 else None

    def messages(self, addr: int):
        """(type, flags, data) of every message of the version 1 object header at `addr`, continuations included."""
        b = self.b
        a = addr + self.base
        if b[a:a + 4] == b"OHDR":
            raise HDF5FormatError("version 2 object header: files written with libver = latest are not supported")
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, a)
        if ver != 1:
            raise HDF5FormatError(f"object header version {ver}")
        blocks = [(a + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                data = b[p + 8:p + 8 + msize]
                if mtype == 0x0010:                                   # continuation: offset, length
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((off + self.base, ln))
                out.append((mtype, mflags, data))
                p += 8 + msize
                left -= 8 + msize
        return out

    def heap_name(self, heap_addr: int, off: int) -> bytes:
        b = self.b
        a = heap_addr + self.base
        if b[a:a + 4] != b"HEAP":
            raise HDF5FormatError("local heap signature")
        _, _, data_addr = struct.unpack_from("<QQQ", b, a + 8)
        s = data_addr + self.base + off
        return b[s:b.find(b"\0", s)]

    def group_entries(self, btree_addr: int, heap_addr: int):
        """{name: object header address} of a symbol-table group."""
        b = self.b
        out = {}
        stack = [btree_addr]
        while stack:
            a = stack.pop() + self.base
            if b[a:a + 4] == b"SNOD":
                nsym = struct.unpack_from("<H", b, a + 6)[0]
                for i in range(nsym):
                    name_off, ohdr = struct.unpack_from("<QQ", b, a + 8 + 40 * i)
                    out[self.heap_name(heap_addr, name_off)] = ohdr
                continue
            if b[a:a + 4] != b"TREE":
                raise HDF5FormatError("group B-tree signature")
            ntype, _, used = struct.unpack_from("<BBH", b, a + 4)
            if ntype != 0:
                raise HDF5FormatError("not a group B-tree")
            for i in range(used):
                stack.append(struct.unpack_from("<Q", b, a + 24 + 8 + 16 * i)[0])     # key, child, key, child, ...
        return out


def _dtype_of(data: bytes) -> np.dtype:
    cls, ver = data[0] & 0x0F, data[0] >> 4
    size = struct.unpack_from("<I", data, 4)[0]
    if ver not in (1, 2, 3) or data[1] & 1:
        raise HDF5FormatError("big-endian or unknown datatype encoding")
    if cls == 1 and size in (4, 8):
        prec, eloc, esz, mloc, msz, bias = struct.unpack_from("<HBBBBI", data, 10)
        if (size, prec, eloc, esz, mloc, msz, bias) == (8, 64, 52, 11, 0, 52, 1023):
            return np.dtype("<f8")
        if (size, prec, eloc, esz, mloc, msz, bias) == (4, 32, 23, 8, 0, 23, 127):
            return np.dtype("<f4")
    if cls == 0 and size in (1, 2, 4, 8):
        return np.dtype(("<i" if data[1] & 0x08 else "<u") + str(size))
    raise HDF5FormatError("unsupported datatype (IEEE little-endian floats and integers only)")


def read_matrix(fname: str, name: str = "A") -> np.ndarray:
    """The Julia array stored as dataset `name` of the root group (Fortran-ordered, float64 for the reference's files)."""
    import mmap
    with open(fname, "rb") as f:
        buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)      # the matrix can be many GB: one copy, at the end
    try:
        return _read_matrix(buf, name)
    finally:
        try:
            buf.close()
        except BufferError:              # a view is still alive (only on an error path)
            pass


def _read_matrix(buf, name: str) -> np.ndarray:
    F = _File(buf)
    if F.root_cache is not None:
        btree, heap = F.root_cache
    else:
        st = [d for t, _, d in F.messages(F.root_header) if t == 0x0011]
        if not st:
            raise HDF5FormatError("root group without a symbol table (link-message groups are not supported)")
        btree, heap = struct.unpack_from("<QQ", st[0], 0)
    entries = F.group_entries(btree, heap)
    if name.encode() not in entries:
        raise KeyError(f'no dataset "{name}" in the root group (found {sorted(k.decode() for k in entries)})')
    dims = dtype = layout = None
    for mtype, _, data in F.messages(entries[name.encode()]):
        if mtype == 0x0001:
            ver, rank, flags = data[0], data[1], data[2]
            off = 8 if ver == 1 else 4
            dims = struct.unpack_from("<%dQ" % rank, data, off)
        elif mtype == 0x0003:
            dtype = _dtype_of(data)
        elif mtype == 0x0008:
            ver = data[0]
            if ver == 3:
                cls = data[1]
                if cls == 1:
                    layout = struct.unpack_from("<QQ", data, 2)
                elif cls == 0:
                    n = struct.unpack_from("<H", data, 2)[0]
                    layout = ("compact", data[4:4 + n])
                else:
                    raise HDF5FormatError("chunked datasets are not supported")
            elif ver in (1, 2):
                rank, cls = data[1], data[2]
                if cls != 1:
                    raise HDF5FormatError("only contiguous version-1/2 layouts are supported")
                addr = struct.unpack_from("<Q", data, 8)[0]
                layout = (addr, None)
            else:
                raise HDF5FormatError(f"data layout message version {ver}")
        elif mtype == 0x000B:
            raise HDF5FormatError("filtered (compressed) datasets are not supported")
    if dims is None or dtype is None or layout is None:
        raise HDF5FormatError("dataset header without dataspace, datatype or layout message")
    count = int(np.prod(dims)) if dims else 1
    if layout[0] == "compact":
        raw = layout[1]
    else:
        addr, size = layout
        if addr == UNDEF:
            return np.zeros(dims[::-1], dtype=dtype, order="F")     # never written: fill value zero
        if addr + F.base + count * dtype.itemsize > len(buf):
            raise HDF5FormatError("dataset extends beyond the end of the file")
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=addr + F.base).reshape(dims)
        out = np.array(a.T, order="F")                                        # Julia's dims and memory order; owns its data
        del a
        return out
    if len(raw) < count * dtype.itemsize:
        raise HDF5FormatError("dataset extends beyond the end of the file")
    return np.array(np.frombuffer(raw, dtype=dtype, count=count).reshape(dims).T, order="F")
