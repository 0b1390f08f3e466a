"""On-disk form of an LsqDB: `<dbpath>_info.json` + `<dbpath>_kron.*`.

Mirrors `/root/reference/src/lsq_db.jl:56-138`: `infofile`/`kronfile` (`:56-58`),
`save_info` (`:96-104`: {"version", "basis", "configs"}), `save_kron`/`load_kron`
(`:111-132`: HDF5 dataset "A"), `_backupfile` (`:73-80`), version gate (`:84-90`).
HDF5 (h5py) is not available in this image, so the matrix is stored as
`<dbpath>_kron.npy` (Fortran-ordered float64, the same bytes as dataset "A");
`_kron.h5` is read/written when h5py can be imported.
"""
from __future__ import annotations

import json
import os
import random
import string

import numpy as np

from .basis import NBodyBasis
from .prototypes import Dat

VERSION = 2


def infofile(dbpath: str) -> str:
    return dbpath + "_info.json"


def kronfile(dbpath: str) -> str:
    try:
        import h5py  # noqa: F401
        return dbpath + "_kron.h5"
    except Exception:
        return dbpath + "_kron.npy"


def _backupfile(fname: str):
    if os.path.exists(fname):
        suffix = "".join(random.choice(string.ascii_letters + string.digits) for _ in range(5))
        os.replace(fname, fname + "." + suffix)


def save(db) -> None:
    if db.dbpath == "":
        raise ValueError("db has no dbpath")
    d = os.path.dirname(db.dbpath)
    if d and not os.path.isdir(d):
        raise FileNotFoundError(f"directory {d} does not exist")
    info = {"version": VERSION, "basis": db.basis.write_dict(), "configs": [c.write_dict() for c in db.configs]}
    _backupfile(infofile(db.dbpath))
    with open(infofile(db.dbpath), "w") as f:
        json.dump(info, f)
    kf = kronfile(db.dbpath)
    _backupfile(kf)
    Psi = np.asfortranarray(db.Psi)
    if kf.endswith(".h5"):
        import h5py
        with h5py.File(kf, "w") as f:
            # Julia's column-major "A" appears with reversed dims to row-major readers
            f.create_dataset("A", data=np.ascontiguousarray(Psi.T))
    else:
        np.save(kf, Psi)


def load_into(db, dbpath: str) -> None:
    with open(infofile(dbpath)) as f:
        info = json.load(f)
    if info.get("version") != VERSION:
        raise RuntimeError(f"This LsqDB was produced with an incompatible version ({info.get('version')}) of the code")
    db.basis = NBodyBasis.read_dict(info["basis"])
    db.configs = [Dat.read_dict(c) for c in info["configs"]]
    db.dbpath = dbpath
    kf = kronfile(dbpath)
    if kf.endswith(".h5"):
        import h5py
        with h5py.File(kf, "r") as f:
            Psi = np.asfortranarray(np.array(f["A"]).T)
    else:
        Psi = np.load(kf)
    db._Psi_host = np.asfortranarray(Psi, dtype=np.float64)
    db._Psi_dev = None
    db._dev_basis = None
    db.nrows = db._Psi_host.shape[0]
