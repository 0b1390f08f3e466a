"""`LsqDB`: the least-squares database (design matrix + configs + basis).

Mirrors `/root/reference/src/lsq_db.jl`:
  LsqDB(dbpath, basis, configs; verbose, maxnthreads)   :171-200
  _alloc_lsq_matrix (row assignment, zeros(nrows, nbasis)) :237-250
  set_matrows!/matrows                                   :230-235
  safe_append! (eval_obs -> vec_obs -> Psi[irows,:] = .) :268-278   <- replaced by ipf_assemble
  flush / save / load                                    :96-138,165-169  (lsq_db_io.py)
and the README-era argument order `LsqDB(fname, configs, basis)` (`README.md:61`).

Psi is device resident (`DeviceMatrix`); `db.Psi` materialises it on the host
lazily.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from . import _capi
from .basis import NBodyBasis, as_basis
from .datatypes import obs_len
from .obsiter import observations
from .prototypes import Dat, ENERGY, FORCES, VIRIAL


def matrows(d: Dat, okey: str) -> np.ndarray:
    """0-based row indices of observation `okey` of `d` (`src/lsq_db.jl:235`, 1-based there)."""
    first, n = d.rows[okey]
    return np.arange(first, first + n)


def set_matrows(d: Dat, okey: str, first: int, n: int) -> Dat:
    d.rows[okey] = (int(first), int(n))
    return d


_OKEY_ID = {ENERGY: 0, FORCES: 1, VIRIAL: 2}


def _alloc_rows(configs: Sequence[Dat], layout: dict = None) -> int:
    """Row assignment of `_alloc_lsq_matrix` (`src/lsq_db.jl:237-250`): prefix sum of the
    observation lengths in `observations(configs)` order.  Returns nrows.  When `layout` is a dict
    it receives the per-observation arrays (config, okey id, first row, length, natoms, configtype id)
    that `collect_observations` and `flatten_configs` use for their vectorised paths."""
    nrows = 0
    flat = []            # (config, okey id, first row, length, natoms, configtype id) per E/F/V observation
    ext = flat.extend
    cts = {}
    plain = True
    okid = _OKEY_ID.get
    for icfg, d in enumerate(configs):      # observations(configs) order: config-major, then D's key order
        ci = cts.setdefault(d.configtype, len(cts))
        na = d.at.X.shape[0]
        rows = {}
        for okey, v in d.D.items():
            n = v.shape[0]
            rows[okey] = (nrows, n)
            k = okid(okey)
            if k is None:
                plain = False
            else:
                ext((icfg, k, nrows, n, na, ci))
            nrows += n
        d.rows = rows
    if layout is not None:
        a = np.asarray(flat, dtype=np.int64).reshape(-1, 6)
        layout.update(nconfigs=len(configs), plain=plain, cts=list(cts.keys()), contiguous=plain,
                      cfg=a[:, 0].copy(), key=a[:, 1].copy(), first=a[:, 2].copy(), len=a[:, 3].copy(),
                      nat=a[:, 4].astype(np.float64), ctid=a[:, 5].copy())
    return nrows


def _flatten_atoms(configs: Sequence[Dat]):
    """Atom data of a configuration list as the flat arrays `ipf_assemble` takes: (atom_off, pos, cell, pbc, Z)."""
    ncfg = len(configs)
    ats = [d.at for d in configs]
    nat = np.array([a.X.shape[0] for a in ats], dtype=np.int64)
    off = np.zeros(ncfg + 1, dtype=np.int64)
    np.cumsum(nat, out=off[1:])
    if ncfg:
        pos = np.ascontiguousarray(np.concatenate([a.X for a in ats], axis=0), dtype=np.float64)
        cell = np.ascontiguousarray(np.concatenate([a.cell for a in ats]), dtype=np.float64).reshape(ncfg, 3, 3)
        pbc = np.concatenate([a.pbc for a in ats]).astype(np.uint8).reshape(ncfg, 3)
        Z = np.ascontiguousarray(np.concatenate([a.Z for a in ats]), dtype=np.int32)
    else:
        pos = np.zeros((0, 3)); cell = np.zeros((0, 3, 3)); pbc = np.zeros((0, 3), dtype=np.uint8)
        Z = np.zeros(0, dtype=np.int32)
    return off, pos, cell, pbc, Z


def _rows_from_layout(layout: dict, ncfg: int):
    """1-based first rows (0 = absent) per observation type from the arrays `_alloc_rows` recorded, with the length
    check of every E / F / V block; None when the layout does not describe this list."""
    if not (layout and layout.get("nconfigs") == ncfg and "cfg" in layout):
        return None
    lc, lk, lf, ll = layout["cfg"], layout["key"], layout["first"], layout["len"]
    expect = np.where(lk == 0, 1, np.where(lk == 1, 3 * layout["nat"].astype(np.int64), 6))
    bad = np.nonzero(ll != expect)[0]
    if bad.size:
        b = int(bad[0])
        k = (ENERGY, FORCES, VIRIAL)[int(lk[b])]
        raise ValueError(f"observation {k} of config {int(lc[b])} has length {int(ll[b])}, "
                         f"expected {int(expect[b])}")
    rows = {}
    for ki, k in enumerate((ENERGY, FORCES, VIRIAL)):
        r = np.zeros(ncfg, dtype=np.int64)
        m = lk == ki
        r[lc[m]] = lf[m] + 1             # 1-based at the ABI, 0 = absent
        rows[k] = r
    return rows


def flatten_configs(configs: Sequence[Dat], layout: dict = None):
    """The flat arrays `ipf_assemble` takes.  `layout` (from `_alloc_rows` on the same list) gives the
    1-based row blocks without another pass over the configurations."""
    ncfg = len(configs)
    off, pos, cell, pbc, Z = _flatten_atoms(configs)
    rows = _rows_from_layout(layout, ncfg)
    if rows is not None:
        return off, pos, cell, pbc, Z, rows
    rows = {}
    for k in (ENERGY, FORCES, VIRIAL):
        r = np.zeros(ncfg, dtype=np.int64)
        for c, d in enumerate(configs):
            if k in d.D:
                if k not in d.rows:
                    raise ValueError("rows not assigned; call _alloc_rows first")
                if d.rows[k][1] != obs_len(k, len(d.at)):
                    raise ValueError(f"observation {k} of config {c} has length {d.rows[k][1]}, "
                                     f"expected {obs_len(k, len(d.at))}")
                r[c] = d.rows[k][0] + 1      # 1-based at the ABI, 0 = absent
        rows[k] = r
    return off, pos, cell, pbc, Z, rows


class LsqDB:
    """`LsqDB(dbpath, basis, configs)` -- also accepts `LsqDB(dbpath, configs, basis)`
    and `LsqDB(dbpath)` (load)."""

    def __init__(self, dbpath: str = "", basis=None, configs=None, *, verbose: bool = False,
                 maxnthreads=None, ctx: _capi.Context = None, _assemble: bool = True):
        if basis is None and configs is None:
            from .lsq_db_io import load_into
            self.ctx = ctx
            load_into(self, dbpath)
            return
        # README-era order LsqDB(fname, configs, basis)
        if _looks_like_configs(basis) and not _looks_like_configs(configs):
            basis, configs = configs, basis
        self.basis: NBodyBasis = as_basis(basis)
        self.configs: List[Dat] = list(configs)
        self.dbpath = dbpath
        self.ctx = ctx
        self._Psi_dev = None
        self._Psi_host = None
        self._dev_basis = None
        self._obs_layout = {}
        self.nrows = _alloc_rows(self.configs, self._obs_layout)
        if _assemble:
            self._assemble()
            if dbpath != "":
                try:
                    self.flush()
                except Exception as e:       # `src/lsq_db.jl:189-194`: warn, keep the data
                    import warnings
                    warnings.warn("something went wrong trying to save the db to disk, but the data "
                                  f"should be ok; if it is crucial to keep it, try to save manually. ({e})")

    # ---- device plumbing
    def _context(self) -> _capi.Context:
        if self.ctx is None:
            self.ctx = _capi.default_context()
        return self.ctx

    def _assemble(self):
        ctx = self._context()
        self._dev_basis = _capi.DeviceBasis(ctx, self.basis)
        self._Psi_dev = _capi.DeviceMatrix(ctx, self.nrows, len(self.basis))
        # no final synchronisation: the solve / download that follows is ordered behind the assembly on the
        # context's stream, and the host-side work of lsqfit overlaps the last kernels.  Large sets go down in two
        # calls so that flattening the second half overlaps the kernels of the first.
        ncfg = len(self.configs)
        rows = _rows_from_layout(self._obs_layout, ncfg)
        parts = [(0, ncfg)] if (rows is None or ncfg < 256) else [(0, ncfg // 2), (ncfg // 2, ncfg)]
        ctx.set_async(True)
        try:
            for a, b in parts:
                if rows is None:
                    off, pos, cell, pbc, Z, r = flatten_configs(self.configs, self._obs_layout)
                else:
                    off, pos, cell, pbc, Z = _flatten_atoms(self.configs[a:b])
                    r = {k: np.ascontiguousarray(v[a:b]) for k, v in rows.items()}
                self._assemble_call(ctx, b - a, off, pos, cell, pbc, Z, r)
        finally:
            ctx.set_async(False)
        self._Psi_host = None

    def _assemble_call(self, ctx, ncfg, off, pos, cell, pbc, Z, rows):
        ctx.check(ctx.lib.ipf_assemble(
            ctx.h, self._dev_basis.h, ncfg, _capi.ptr(off, C.c_int64),
            _capi.ptr(pos, C.c_double), _capi.ptr(cell, C.c_double), _capi.ptr(pbc, C.c_uint8),
            _capi.ptr(Z, C.c_int32), _capi.ptr(rows[ENERGY], C.c_int64), _capi.ptr(rows[FORCES], C.c_int64),
            _capi.ptr(rows[VIRIAL], C.c_int64), self._Psi_dev.h))

    @property
    def Psi_dev(self) -> _capi.DeviceMatrix:
        if self._Psi_dev is None:
            if self._Psi_host is None:
                raise RuntimeError("database has been deleted (deldb)")
            ctx = self._context()
            self._Psi_dev = _capi.DeviceMatrix(ctx, *self._Psi_host.shape)
            self._Psi_dev.from_host(self._Psi_host)
        return self._Psi_dev

    @property
    def Psi(self) -> np.ndarray:
        """Host copy of the design matrix (column-major), materialised lazily."""
        if self._Psi_host is None:
            if self._Psi_dev is None:
                return np.zeros((0, 0), order="F")
            self._Psi_host = self._Psi_dev.to_host()
        return self._Psi_host

    def delete(self):
        """`deldb=true` (`src/lsq.jl:442-446`)."""
        if self._Psi_dev is not None:
            self._Psi_dev.close()
        self._Psi_dev = None
        self._Psi_host = None
        self.dbpath = ""

    def flush(self):
        from .lsq_db_io import save
        save(self)

    def __len__(self):
        return len(self.configs)


def _looks_like_configs(x) -> bool:
    return isinstance(x, (list, tuple)) and len(x) > 0 and isinstance(x[0], Dat)


def dbpath(db: LsqDB) -> str:
    return db.dbpath


def filter_basis(db: LsqDB, pred) -> np.ndarray:
    """`Ib2 = filter_basis(db, b -> bodyorder(b) < 3)` (`README.md:92-96`): 0-based column indices."""
    return np.array([k for k, b in enumerate(db.basis) if pred(b)], dtype=np.int64)


def filter_configs(db: LsqDB, pred) -> np.ndarray:
    """0-based config indices with pred(dat) true (`README.md:92-98`)."""
    return np.array([k for k, d in enumerate(db.configs) if pred(d)], dtype=np.int64)
