"""On-disk form of an LsqDB: `<dbpath>_info.json` + `<dbpath>_kron.*`.

Mirrors `/root/reference/src/lsq_db.jl:56-138`: `infofile`/`kronfile` (`:56-58`),
`save_info` (`:96-104`: {"version", "basis", "configs"}), `save_kron`/`load_kron`
(`:111-132`: HDF5 dataset "A"), `_backupfile` (`:73-80`), version gate (`:84-90`).
The matrix goes to `<dbpath>_kron.h5`, dataset "A", as in the reference: through h5py when it can be imported, else
through the minimal writer / reader of `hdf5_min.py` (version 0 superblock, one contiguous float64 dataset; written
from the format specification -- no HDF5 library exists in this image to check it against).  `<dbpath>_kron.npy`
(Fortran-ordered float64, what earlier versions of this package wrote) is still read when no `.h5` file exists.
"""
from __future__ import annotations

import json
import os
import random
import string

import numpy as np

from . import hdf5_min
from .basis import NBodyBasis
from .prototypes import Dat

VERSION = 2


def infofile(dbpath: str) -> str:
    return dbpath + "_info.json"


def kronfile(dbpath: str) -> str:
    return dbpath + "_kron.h5"


def _have_h5py() -> bool:
    try:
        import h5py  # noqa: F401
        return True
    except Exception:
        return False


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
    if _have_h5py():
        import h5py
        with h5py.File(kf, "w") as f:
            # Julia's column-major "A" appears with reversed dims to row-major readers
            f.create_dataset("A", data=np.ascontiguousarray(Psi.T))
    else:
        hdf5_min.write_matrix(kf, Psi, "A")


def load_into(db, dbpath: str) -> None:
    with open(infofile(dbpath)) as f:
        info = json.load(f)
    if info.get("version") != VERSION:
        raise RuntimeError(f"This LsqDB was produced with an incompatible version ({info.get('version')}) of the code")
    db.basis = NBodyBasis.read_dict(info["basis"])
    db.configs = [Dat.read_dict(c) for c in info["configs"]]
    db.dbpath = dbpath
    kf = kronfile(dbpath)
    if not os.path.exists(kf) and os.path.exists(dbpath + "_kron.npy"):
        Psi = np.load(dbpath + "_kron.npy")
    elif _have_h5py():
        import h5py
        with h5py.File(kf, "r") as f:
            Psi = np.asfortranarray(np.array(f["A"]).T)
    else:
        Psi = hdf5_min.read_matrix(kf, "A")
    db._Psi_host = np.asfortranarray(Psi, dtype=np.float64)
    db._Psi_dev = None
    db._dev_basis = None
    db.nrows = db._Psi_host.shape[0]
