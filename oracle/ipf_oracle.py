"""ctypes wrapper around oracle/libipf_oracle.so (TEST INFRASTRUCTURE ONLY --
see the header of ipf_oracle.c; parity unpinned).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libipf_oracle.so")
    src = os.path.join(_HERE, "ipf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip, i64p, u8p = (C.POINTER(C.c_double), C.POINTER(C.c_int32),
                             C.POINTER(C.c_int64), C.POINTER(C.c_uint8))
        L.ipfo_neighbours.restype = C.c_int64
        L.ipfo_neighbours.argtypes = [C.c_int32, dp, dp, u8p, C.c_double, C.c_int64, ip, ip, ip, dp]
        L.ipfo_assemble_block.restype = C.c_int32
        L.ipfo_assemble_block.argtypes = [C.c_int32, C.c_int32, C.c_int32, ip, u8p, C.c_int32,
                                          C.c_double, C.c_double, C.c_double, C.c_double,
                                          C.c_int64, i64p, dp, dp, u8p, i64p, i64p, i64p,
                                          C.c_int64, dp, C.c_int64, C.c_int32, C.c_int32]
        L.ipfo_assemble_block_z.restype = C.c_int32
        L.ipfo_assemble_block_z.argtypes = L.ipfo_assemble_block.argtypes + [ip, C.c_int32, ip, C.c_int32, C.c_double, C.c_double]
        L.ipfo_eval_config.restype = C.c_int32
        L.ipfo_eval_config.argtypes = [C.c_int32, C.c_int32, C.c_int32, ip, u8p, C.c_int32,
                                       C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.c_int32, dp, dp, u8p, dp, dp, dp]
        L.ipfo_eval_config_z.restype = C.c_int32
        L.ipfo_eval_config_z.argtypes = L.ipfo_eval_config.argtypes + [ip, C.c_int32, ip, C.c_int32, C.c_double, C.c_double]
        L.ipfo_num_threads.restype = C.c_int32
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def neighbours(X, cell, pbc, rcut):
    """Canonical neighbour list: (I, J, S[n,3], R[n,3]) sorted by (i, j, Sx, Sy, Sz)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    cell = np.ascontiguousarray(cell, dtype=np.float64)
    pbc = np.ascontiguousarray(pbc, dtype=np.uint8)
    L = lib()
    n = L.ipfo_neighbours(len(X), _p(X, C.c_double), _p(cell, C.c_double), _p(pbc, C.c_uint8),
                          float(rcut), 0, None, None, None, None)
    I = np.empty(n, dtype=np.int32); J = np.empty(n, dtype=np.int32)
    S = np.empty((n, 3), dtype=np.int32); R = np.empty((n, 3), dtype=np.float64)
    L.ipfo_neighbours(len(X), _p(X, C.c_double), _p(cell, C.c_double), _p(pbc, C.c_uint8),
                      float(rcut), n, _p(I, C.c_int32), _p(J, C.c_int32), _p(S, C.c_int32),
                      _p(R, C.c_double))
    return I, J, S, R


def eval_config(block, at):
    """Dense (E[nb], F[nb, 3 nat], V[nb, 3, 3]) of one basis block on one Atoms."""
    kind, a, r0, rc1, rc2 = block.desc.params()
    nat = len(at)
    E = np.zeros(block.nb); F = np.zeros((block.nb, 3 * nat)); V = np.zeros((block.nb, 9))
    mb = np.ascontiguousarray(block.mono_b, dtype=np.int32)
    me = np.ascontiguousarray(block.mono_exp, dtype=np.uint8)
    pbc = np.ascontiguousarray(at.pbc, dtype=np.uint8)
    X = np.ascontiguousarray(at.X); cell = np.ascontiguousarray(at.cell)
    Z = np.ascontiguousarray(at.Z, dtype=np.int32)
    zs = np.zeros(4, dtype=np.int32)
    zs[:len(block.zs)] = block.zs
    rc = lib().ipfo_eval_config_z(block.N, block.nb, len(mb), _p(mb, C.c_int32), _p(me, C.c_uint8),
                                  kind, a, r0, rc1, rc2, nat, _p(X, C.c_double),
                                  _p(cell, C.c_double), _p(pbc, C.c_uint8),
                                  _p(E, C.c_double), _p(F, C.c_double), _p(V, C.c_double),
                                  _p(Z, C.c_int32), int(block.zc), _p(zs, C.c_int32),
                                  int(block.env_t), float(block.env_rc[0]), float(block.env_rc[1]))
    assert rc == 0
    return E, F, V.reshape(-1, 3, 3)


def flatten_configs(configs):
    """Flat arrays the C boundary takes (shared shape with the product's C-ABI)."""
    nat = np.array([len(d.at) for d in configs], dtype=np.int64)
    off = np.zeros(len(configs) + 1, dtype=np.int64)
    np.cumsum(nat, out=off[1:])
    pos = np.ascontiguousarray(np.concatenate([d.at.X for d in configs], axis=0))
    cell = np.ascontiguousarray(np.stack([d.at.cell for d in configs]))
    pbc = np.ascontiguousarray(np.stack([d.at.pbc for d in configs]).astype(np.uint8))
    return off, pos, cell, pbc


def assemble(basis, configs, rowE, rowF, rowV, nrows, mode=0, nthreads=0):
    """Column-major Psi[nrows, len(basis)] (returned as a Fortran-ordered array).
    rowE/F/V: 1-based first rows, 0 = absent."""
    off, pos, cell, pbc = flatten_configs(configs)
    Psi = np.zeros((nrows, len(basis)), dtype=np.float64, order="F")
    rowE = np.ascontiguousarray(rowE, dtype=np.int64)
    rowF = np.ascontiguousarray(rowF, dtype=np.int64)
    rowV = np.ascontiguousarray(rowV, dtype=np.int64)
    L = lib()
    Zall = np.ascontiguousarray(np.concatenate([np.asarray(d.at.Z, dtype=np.int32) for d in configs])
                                if len(configs) else np.zeros(0, dtype=np.int32))
    for blk in basis.blocks:
        kind, a, r0, rc1, rc2 = blk.desc.params()
        mb = np.ascontiguousarray(blk.mono_b, dtype=np.int32)
        me = np.ascontiguousarray(blk.mono_exp, dtype=np.uint8)
        zs = np.zeros(4, dtype=np.int32)
        zs[:len(blk.zs)] = blk.zs
        rc = L.ipfo_assemble_block_z(blk.N, blk.nb, len(mb), _p(mb, C.c_int32), _p(me, C.c_uint8),
                                     kind, a, r0, rc1, rc2, len(configs), _p(off, C.c_int64),
                                     _p(pos, C.c_double), _p(cell, C.c_double), _p(pbc, C.c_uint8),
                                     _p(rowE, C.c_int64), _p(rowF, C.c_int64), _p(rowV, C.c_int64),
                                     blk.col0, _p(Psi, C.c_double), nrows, mode, nthreads,
                                     _p(Zall, C.c_int32), int(blk.zc), _p(zs, C.c_int32),
                                     int(blk.env_t), float(blk.env_rc[0]), float(blk.env_rc[1]))
        assert rc == 0
    return Psi


def num_threads() -> int:
    return int(lib().ipfo_num_threads())
