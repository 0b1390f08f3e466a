"""`<dbpath>_kron.h5` through the minimal HDF5 writer / reader (`/root/reference/src/lsq_db.jl:111-126`: dataset "A").

No HDF5 library exists in the image: the checks are the round trip, the structures of the written file against the HDF5
File Format Specification byte by byte (the test restates the offsets independently of the writer's code), and the
reader's error behaviour.  When h5py can be imported the file is also read with it."""
import os
import struct

import numpy as np
import pytest

import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import hdf5_min as h5
from ipfitting_jl_b200 import synth
from ipfitting_jl_b200.lsq_db import LsqDB

UNDEF = 0xFFFFFFFFFFFFFFFF


def _write(tmp_path, A, name="A"):
    f = str(tmp_path / "m_kron.h5")
    h5.write_matrix(f, A, name)
    return f, open(f, "rb").read()


@pytest.mark.parametrize("shape", [(7, 5), (1, 1), (1, 9), (13, 1), (0, 4), (300, 17)])
def test_roundtrip(tmp_path, shape):
    A = np.random.default_rng(3).standard_normal(shape)
    f, _ = _write(tmp_path, A)
    B = h5.read_matrix(f)
    assert B.shape == A.shape and B.dtype == np.float64 and B.flags["F_CONTIGUOUS"]
    assert np.array_equal(A.view(np.int64) if A.size else A, B.view(np.int64) if B.size else B)        # bit-exact, NaN-safe


def test_special_values_bit_exact(tmp_path):
    A = np.array([[np.nan, -0.0, np.inf], [5e-324, -np.inf, 1.7976931348623157e308]])
    f, _ = _write(tmp_path, A)
    assert np.array_equal(h5.read_matrix(f).view(np.int64), np.asfortranarray(A).view(np.int64))


def test_file_structure_follows_the_specification(tmp_path):
    A = np.arange(12, dtype=np.float64).reshape(3, 4)          # Julia 3 x 4
    _, b = _write(tmp_path, A)
    # ---- superblock, version 0 (spec III.A)
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    assert b[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])          # versions 0; 8-byte offsets and lengths
    leaf_k, internal_k, flags = struct.unpack_from("<HHI", b, 16)
    assert (leaf_k, internal_k, flags) == (4, 16, 0)
    base, free, eof, driver = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == UNDEF and driver == UNDEF and eof == len(b)
    name_off, root_hdr, cache, _ = struct.unpack_from("<QQII", b, 56)           # root symbol-table entry
    btree, heap = struct.unpack_from("<QQ", b, 80)
    assert name_off == 0 and cache == 1 and root_hdr == 96
    # ---- root object header, version 1 (spec IV.A.1.a): symbol-table message 0x0011 with the same addresses
    ver, _, nmsg, refc, hsize = struct.unpack_from("<BBHII", b, root_hdr)
    assert (ver, nmsg, refc) == (1, 1, 1) and root_hdr % 8 == 0
    mtype, msize, _ = struct.unpack_from("<HHB", b, root_hdr + 16)
    assert mtype == 0x0011 and msize == 16 and hsize == 8 + msize
    assert struct.unpack_from("<QQ", b, root_hdr + 24) == (btree, heap)
    # ---- group B-tree node (spec III.A.1): type 0, leaf level, one entry, no siblings, full node allocated
    assert b[btree:btree + 4] == b"TREE"
    ntype, level, used, left, right = struct.unpack_from("<BBHQQ", b, btree + 4)
    assert (ntype, level, used, left, right) == (0, 0, 1, UNDEF, UNDEF)
    key0, child0, key1 = struct.unpack_from("<QQQ", b, btree + 24)
    node_size = 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8
    assert heap == btree + node_size
    # ---- local heap (spec III.D): data segment right behind the header, no free block
    assert b[heap:heap + 4] == b"HEAP" and b[heap + 4] == 0
    seg_size, free_head, seg = struct.unpack_from("<QQQ", b, heap + 8)
    assert seg == heap + 32 and seg_size % 8 == 0 and free_head == 1
    assert b[seg + key0] == 0                                                   # key 0: the empty string
    assert b[seg + key1:seg + key1 + 2] == b"A\0"                               # key 1: the largest name of child 0
    # ---- symbol-table node (spec III.B): one entry pointing at the dataset header
    assert b[child0:child0 + 4] == b"SNOD" and b[child0 + 4] == 1
    assert struct.unpack_from("<H", b, child0 + 6)[0] == 1
    e_name, dset, e_cache, _ = struct.unpack_from("<QQII", b, child0 + 8)
    assert e_name == key1 and e_cache == 0 and dset == child0 + 8 + 2 * leaf_k * 40 and dset % 8 == 0
    # ---- dataset object header: dataspace (1), datatype (3), layout (8); every message 8-byte aligned
    ver, _, nmsg, refc, hsize = struct.unpack_from("<BBHII", b, dset)
    assert (ver, nmsg, refc) == (1, 3, 1)
    p, seen = dset + 16, {}
    for _ in range(nmsg):
        mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
        assert msize % 8 == 0
        seen[mtype] = b[p + 8:p + 8 + msize]
        p += 8 + msize
    assert p == dset + 16 + hsize and set(seen) == {1, 3, 8}
    sp = seen[1]
    assert sp[0] == 1 and sp[1] == 2 and sp[2] == 0                             # version 1, rank 2, no maximum dims
    assert struct.unpack_from("<QQ", sp, 8) == (4, 3)                           # Julia dims reversed
    # IEEE binary64 little endian, the byte pattern of H5T_IEEE_F64LE
    assert seen[3][:20] == bytes.fromhex("11203f0008000000" "00004000340b0034ff030000")
    lay = seen[8]
    assert lay[0] == 3 and lay[1] == 1                                          # version 3, contiguous
    addr, size = struct.unpack_from("<QQ", lay, 2)
    assert size == 96 and addr + size == eof and addr % 8 == 0
    # column-major bytes of the Julia matrix = row-major over the reversed dims
    assert np.array_equal(np.frombuffer(b[addr:addr + size], "<f8").reshape(4, 3), A.T)


def test_reader_errors(tmp_path):
    A = np.ones((2, 3))
    f, b = _write(tmp_path, A)
    with pytest.raises(KeyError):
        h5.read_matrix(f, "B")
    bad = str(tmp_path / "bad.h5")
    open(bad, "wb").write(b"not hdf5" + b[8:])
    with pytest.raises(h5.HDF5FormatError):
        h5.read_matrix(bad)
    open(bad, "wb").write(b[:-8])                                               # truncated
    with pytest.raises(h5.HDF5FormatError):
        h5.read_matrix(bad)
    v2 = bytearray(b); v2[8] = 2                                                # superblock of libver = latest
    open(bad, "wb").write(bytes(v2))
    with pytest.raises(h5.HDF5FormatError, match="libver"):
        h5.read_matrix(bad)
    chunked = bytearray(b)
    i = chunked.index(struct.pack("<HHB3x", 0x0008, 24, 0)) + 8
    chunked[i + 1] = 2                                                          # layout class 2 = chunked
    open(bad, "wb").write(bytes(chunked))
    with pytest.raises(h5.HDF5FormatError, match="chunked"):
        h5.read_matrix(bad)


def test_reader_follows_continuations_and_uncached_root(tmp_path):
    """Files of libhdf5 carry more than the minimum: a root entry without cached addresses, object-header continuation
    blocks, NIL messages.  Rebuild the written file that way and read it back."""
    A = np.random.default_rng(5).standard_normal((4, 6))
    _, b = _write(tmp_path, A)
    b = bytearray(b)
    struct.pack_into("<I", b, 72, 0)                                            # root entry: cache type 0
    b[80:96] = b"\0" * 16
    dset = 1056
    # move the layout message into a continuation block at the end of the metadata area, replaced by a continuation
    # message (0x0010: offset, length); a NIL message (type 0) pads the block
    p = dset + 16 + 32 + 32
    layout_msg = bytes(b[p:p + 32])
    cont_at = 1600
    block = struct.pack("<HHB3x", 0, 8, 0) + b"\0" * 8 + layout_msg
    b[cont_at:cont_at + len(block)] = block
    b[p:p + 32] = struct.pack("<HHB3x", 0x0010, 16, 0) + struct.pack("<QQ", cont_at, len(block)) + b"\0" * 8
    # messages: dataspace, datatype, continuation, the 8 bytes left over in the first block (a NIL message of size 0),
    # NIL, layout
    struct.pack_into("<H", b, dset + 2, 6)
    f = str(tmp_path / "c.h5")
    open(f, "wb").write(bytes(b))
    assert np.array_equal(h5.read_matrix(f), A)


def test_lsqdb_writes_and_reads_kron_h5(tmp_path):
    """`test/test_lsq_db.jl:46-75`: the database files are `<path>_info.json` and `<path>_kron.h5`."""
    cfgs = synth.rattled_configs("Si", 1, 2, seed=9, calc=synth.StillingerWeber())
    r0 = ipf.rnn("Si")
    B = ipf.nbpolys(2, ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(2 * r0 - 1, 2 * r0)), 3)
    db = LsqDB(str(tmp_path / "db"), B, cfgs, _assemble=False)
    db._Psi_host = np.asfortranarray(np.random.default_rng(2).random((db.nrows, len(B))))
    db.flush()
    assert sorted(os.listdir(tmp_path)) == ["db_info.json", "db_kron.h5"]
    db2 = LsqDB(str(tmp_path / "db"))
    assert np.array_equal(db2.Psi, db.Psi) and db2.Psi.flags["F_CONTIGUOUS"]
    assert np.array_equal(h5.read_matrix(str(tmp_path / "db_kron.h5")), db.Psi)
    try:
        import h5py
    except Exception:
        return
    with h5py.File(str(tmp_path / "db_kron.h5"), "r") as f:
        assert np.array_equal(np.array(f["A"]).T, db.Psi)


def test_legacy_npy_is_still_read(tmp_path):
    cfgs = synth.rattled_configs("Si", 1, 1, seed=9, calc=synth.StillingerWeber())
    r0 = ipf.rnn("Si")
    B = ipf.nbpolys(2, ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(2 * r0 - 1, 2 * r0)), 3)
    db = LsqDB(str(tmp_path / "db"), B, cfgs, _assemble=False)
    db._Psi_host = np.asfortranarray(np.random.default_rng(2).random((db.nrows, len(B))))
    db.flush()
    os.remove(str(tmp_path / "db_kron.h5"))
    np.save(str(tmp_path / "db_kron.npy"), db.Psi)
    assert np.array_equal(LsqDB(str(tmp_path / "db")).Psi, db.Psi)
