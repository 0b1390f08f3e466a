"""Observation-type plugins E / F / V: vectorisation and weight hooks.

Mirrors `/root/reference/src/datatypes.jl`:
  E: vec_obs :25-26, devec :27, weighthook 1/sqrt(N) :29, err_weighthook 1/N :30
  F: vec_obs :37-38 (+ basis-wide :42-50), devec :39           (weight hooks default 1.0,
                                                             src/prototypes.jl:125,133)
  V: Voigt _IV = [1,5,9,6,3,2] :57, vec_obs :61-62, devec :63-64, hooks :66-67

`eval_obs(okey, B, dat)` for a *basis* B is exactly what the CUDA assembly
replaces (`src/datatypes.jl:28,40,65`); it is not provided per observation here --
the whole-DB entry point is `LsqDB(...)` in lsq_db.py.
"""
from __future__ import annotations

import math

import numpy as np

from .prototypes import ENERGY, FORCES, VIRIAL

# column-major linear indices (0-based) of the 3x3 virial picked by the Voigt map
_IV = (0, 4, 8, 5, 2, 1)


def vec_obs(okey: str, x) -> np.ndarray:
    if okey == ENERGY:
        return np.atleast_1d(np.asarray(x, dtype=np.float64)).reshape(-1).copy()
    if okey == FORCES:
        # [nat,3] -> (x1,y1,z1,x2,...)  (`mat(F)[:]` of a 3 x nat matrix)
        return np.ascontiguousarray(x, dtype=np.float64).reshape(-1).copy()
    if okey == VIRIAL:
        a = np.asarray(x, dtype=np.float64)
        if a.size == 6:
            return a.reshape(6).copy()
        if a.shape != (3, 3):
            raise ValueError("virial must be 3x3 or a Voigt 6-vector")
        return a.reshape(-1, order="F")[list(_IV)].copy()
    return np.asarray(x, dtype=np.float64).reshape(-1).copy()


def devec_obs(okey: str, x):
    x = np.asarray(x, dtype=np.float64)
    if okey == ENERGY:
        if x.size != 1:
            raise ValueError("energy observation must have length 1")
        return float(x.reshape(-1)[0])
    if okey == FORCES:
        return x.reshape(-1, 3)
    if okey == VIRIAL:
        return np.array([[x[0], x[5], x[4]], [x[5], x[1], x[3]], [x[4], x[3], x[2]]])
    return x


def obs_len(okey: str, nat: int) -> int:
    return {ENERGY: 1, FORCES: 3 * nat, VIRIAL: 6}[okey]


def weighthook(okey: str, d) -> float:
    if okey in (ENERGY, VIRIAL):
        return 1.0 / math.sqrt(len(d.at))
    return 1.0


def err_weighthook(okey: str, d) -> float:
    if okey in (ENERGY, VIRIAL):
        return 1.0 / len(d.at)
    return 1.0
