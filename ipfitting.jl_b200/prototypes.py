"""`Atoms`, `Dat`, `LsqDB` data model and the observation hook prototypes.

Mirrors `/root/reference/src/prototypes.jl`:
  Dat            :19-25  (at, configtype, D, rows, info), ctor :34-45, == :28-32
  write_dict/Dat :47-65
  LsqDB struct   :82-87  (basis, configs, Psi, dbpath) -- defined in lsq_db.py
  hook prototypes:106-134 (vec_obs, devec_obs, eval_obs, weighthook, err_weighthook)
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

ENERGY, FORCES, VIRIAL = "E", "F", "V"


class Atoms:
    """Minimal stand-in for `JuLIP.Atoms`: positions X[nat,3], cell[3,3] with
    rows = lattice vectors, pbc[3], atomic numbers Z[nat]."""

    __slots__ = ("X", "cell", "pbc", "Z")

    def __init__(self, X, cell, pbc=(True, True, True), Z=None):
        self.X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 3)
        self.cell = np.ascontiguousarray(cell, dtype=np.float64).reshape(3, 3)
        self.pbc = np.asarray(pbc, dtype=bool).reshape(3)
        n = self.X.shape[0]
        self.Z = (np.zeros(n, dtype=np.int32) if Z is None
                  else np.ascontiguousarray(Z, dtype=np.int32).reshape(n))

    def __len__(self):
        return self.X.shape[0]

    def copy(self):
        return Atoms(self.X.copy(), self.cell.copy(), self.pbc.copy(), self.Z.copy())

    def __eq__(self, o):
        return (isinstance(o, Atoms) and np.array_equal(self.X, o.X)
                and np.array_equal(self.cell, o.cell) and np.array_equal(self.pbc, o.pbc)
                and np.array_equal(self.Z, o.Z))

    def write_dict(self):
        return {"__id__": "JuLIP_Atoms", "X": self.X.tolist(), "cell": self.cell.tolist(),
                "pbc": self.pbc.tolist(), "Z": self.Z.tolist()}

    @staticmethod
    def read_dict(D):
        return Atoms(D["X"], D["cell"], D["pbc"], D["Z"])


class Dat:
    """One configuration + observations (`src/prototypes.jl:19-25`).

    D[okey]    : flat float64 vector (vec_obs form)
    rows[okey] : (first, len) 0-based half-open block in the LSQ matrix; the
                 reference stores the explicit 1-based index vector
                 (`src/lsq_db.jl:244-246`), always contiguous.
    info       : free-form dict ("W", "Vref", fit results, ...)
    """

    def __init__(self, at: Atoms, config_type: str, *, E=None, F=None, V=None,
                 D: Optional[Dict[str, np.ndarray]] = None, info: Optional[dict] = None,
                 okey_order=None):
        # local import: datatypes imports this module for the key constants
        from .datatypes import vec_obs
        self.at = at
        self.configtype = str(config_type)
        self.D: Dict[str, np.ndarray] = {}
        self.rows: Dict[str, tuple] = {}
        self.info: dict = {} if info is None else info
        if D is not None:
            for k, v in D.items():
                self.D[k] = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
        # ctor keywords (`src/prototypes.jl:34-45`)
        given = {ENERGY: E, FORCES: F, VIRIAL: V}
        order = okey_order if okey_order is not None else (ENERGY, FORCES, VIRIAL)
        for k in order:
            if given.get(k) is not None:
                self.D[k] = vec_obs(k, given[k])

    def __len__(self):
        return len(self.at)

    # `==` compares configtype, D, info and the Atoms; NOT rows (`src/prototypes.jl:28-32`)
    def __eq__(self, o):
        if not isinstance(o, Dat):
            return False
        if self.configtype != o.configtype or self.at != o.at:
            return False
        if self.D.keys() != o.D.keys():
            return False
        if any(not np.array_equal(self.D[k], o.D[k]) for k in self.D):
            return False
        return _info_eq(self.info, o.info)

    def write_dict(self):
        return {"__id__": "IPFitting.Dat", "at": self.at.write_dict(),
                "configtype": self.configtype,
                "D": {k: v.tolist() for k, v in self.D.items()},
                # the explicit 1-based index vector, as `set_matrows!(d, okey, collect(nrows .+ (1:len)))` stores it
                # (`src/lsq_db.jl:244-246`, written as is by `write_dict`, `src/prototypes.jl:47-53`)
                "rows": {k: list(range(int(r[0]) + 1, int(r[0]) + int(r[1]) + 1)) for k, r in self.rows.items()},
                "info": _jsonable(self.info)}

    @staticmethod
    def read_dict(Dd):
        d = Dat(Atoms.read_dict(Dd["at"]), Dd["configtype"],
                D={k: np.asarray(v, dtype=np.float64) for k, v in Dd["D"].items()},
                info=Dd.get("info", {}))
        for k, r in Dd.get("rows", {}).items():
            r = [int(v) for v in np.atleast_1d(r)]
            if not r:
                continue
            if any(r[i + 1] != r[i] + 1 for i in range(len(r) - 1)):
                raise ValueError(f"rows[{k!r}] is not a contiguous block: the design matrix stores one block per observation")
            d.rows[k] = (r[0] - 1, len(r))
        return d


def _jsonable(x):
    if isinstance(x, dict):
        return {str(k): _jsonable(v) for k, v in x.items()}
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    return x


def _info_eq(a, b):
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(_info_eq(a[k], b[k]) for k in a)
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.array_equal(np.asarray(a), np.asarray(b))
    return a == b


def configtype(d: Dat) -> str:
    return d.configtype


def observation(d: Dat, okey: str) -> np.ndarray:
    return d.D[okey]


def hasobservation(d: Dat, okey: str) -> bool:
    return okey in d.D


def energy(d: Dat):
    from .datatypes import devec_obs
    return devec_obs(ENERGY, d.D[ENERGY])


def forces(d: Dat):
    from .datatypes import devec_obs
    return devec_obs(FORCES, d.D[FORCES])


def virial(d: Dat):
    from .datatypes import devec_obs
    return devec_obs(VIRIAL, d.D[VIRIAL])
