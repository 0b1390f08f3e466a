"""Calculators at the edge of the hot path: `OneBody`, `combine(basis, c)`, `SumIP`.

The reference takes these from JuLIP (`JuLIP.Potentials.OneBody`, `JuLIP.MLIPs.combine`,
`JuLIP.MLIPs.SumIP`; call sites `src/lsq.jl:29-30,603-606`).  They are restated here
only as far as `lsqfit`/`add_fits!` need them: evaluate E / F / V observations.
"""
from __future__ import annotations

from typing import Dict, Union

import numpy as np

from .datatypes import vec_obs
from .prototypes import ENERGY, FORCES, VIRIAL


class OneBody:
    """`OneBody(E0)` or `OneBody(:Si => E0, ...)`: E = sum_atoms E0[Z]; zero forces and virial."""

    def __init__(self, E0: Union[float, Dict[int, float]] = 0.0):
        self.E0 = E0

    def _e(self, at) -> float:
        if isinstance(self.E0, dict):
            return float(sum(self.E0.get(int(z), 0.0) for z in at.Z))
        return float(self.E0) * len(at)

    def energy(self, at):
        return self._e(at)

    def forces(self, at):
        return np.zeros((len(at), 3))

    def virial(self, at):
        return np.zeros((3, 3))


class NBodyIP:
    """`combine(basis, c)`: the fitted potential sum_b c_b B_b.  Evaluation goes through the
    CUDA assembly (one config batch -> Psi, then Psi c)."""

    def __init__(self, basis, c):
        self.basis = basis
        self.c = np.asarray(c, dtype=np.float64).copy()
        if len(self.c) != len(basis):
            raise ValueError("coefficient vector does not match the basis")

    def _efv(self, at, ctx=None):
        from .lsq_db import LsqDB
        from .prototypes import Dat
        nat = len(at)
        d = Dat(at, "eval", E=0.0, F=np.zeros((nat, 3)), V=np.zeros(6))
        db = LsqDB("", self.basis, [d], ctx=ctx)
        y = db.Psi @ self.c
        out = {k: y[d.rows[k][0]: d.rows[k][0] + d.rows[k][1]] for k in (ENERGY, FORCES, VIRIAL)}
        return out

    def energy(self, at, ctx=None):
        return float(self._efv(at, ctx)[ENERGY][0])

    def forces(self, at, ctx=None):
        return self._efv(at, ctx)[FORCES].reshape(-1, 3)

    def virial(self, at, ctx=None):
        from .datatypes import devec_obs
        return devec_obs(VIRIAL, self._efv(at, ctx)[VIRIAL])


class SumIP:
    """`SumIP(Vref, IP)` (`src/lsq.jl:604-606`)."""

    def __init__(self, *components):
        self.components = list(components)

    def energy(self, at):
        return sum(c.energy(at) for c in self.components)

    def forces(self, at):
        return sum(c.forces(at) for c in self.components)

    def virial(self, at):
        return sum(c.virial(at) for c in self.components)


def linear_part(IP):
    """(NBodyIP, [other components]) of a fitted potential."""
    if isinstance(IP, NBodyIP):
        return IP, []
    if isinstance(IP, SumIP):
        lin = [c for c in IP.components if isinstance(c, NBodyIP)]
        if len(lin) == 1:
            return lin[0], [c for c in IP.components if c is not lin[0]]
    return None, [IP]


def vref_observation(V, okey: str, dat) -> np.ndarray:
    """`vec_obs(okey, eval_obs(okey, Vref, dat))` (`src/lsq.jl:98`) for a reference calculator."""
    if okey == ENERGY:
        return vec_obs(ENERGY, V.energy(dat.at))
    if okey == FORCES:
        return vec_obs(FORCES, V.forces(dat.at))
    if okey == VIRIAL:
        return vec_obs(VIRIAL, V.virial(dat.at))
    raise KeyError(okey)
