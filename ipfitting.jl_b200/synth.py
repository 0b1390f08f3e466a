"""Deterministic synthetic rattled-crystal datasets with analytic labels.

The reference's tests generate their data in-process
(`/root/reference/test/test_fit2b.jl:12-22`, `test/test_fit3b.jl:27-35`):
`bulk(species; cubic=true, pbc=true) * L`, `rattle!(at, rand()*rmax)`, labels
from a cheap analytic calculator (LennardJones*C2Shift, StillingerWeber).
Julia's RNG stream and JuLIP's `rattle!` cannot be reproduced here, so the
recipe is restated and frozen:

  * amplitude A = u * rmax, u ~ U(0,1) per config; every coordinate += A * U(-1,1)
  * symmetric cell strain I + eps, eps_ab ~ U(-strain, strain), applied to cell and positions
  * counter-based RNG: numpy Philox, key = seed

Label calculators are plain numpy (host side data generation, not the hot path).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from .prototypes import Atoms, Dat

_Z = {"Si": 14, "Cu": 29, "W": 74, "C": 6}
_A0 = {"Si": 5.43, "Cu": 3.61, "W": 3.165, "C": 3.567}


def bulk(species: str, cubic: bool = True) -> Atoms:
    """Conventional cubic cell of the ground-state crystal (JuLIP `bulk(...; cubic=true)`)."""
    a = _A0[species]
    if species in ("Si", "C"):
        f = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
        frac = np.vstack([f, f + 0.25])
    elif species == "Cu":
        frac = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    elif species == "W":
        frac = np.array([[0, 0, 0], [.5, .5, .5]])
    else:
        raise ValueError(species)
    return Atoms(frac * a, np.eye(3) * a, (True, True, True),
                 np.full(len(frac), _Z[species], dtype=np.int32))


def zincblende(a: float, za: int, zb: int) -> Atoms:
    f = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    frac = np.vstack([f, f + 0.25])
    Z = np.array([za] * 4 + [zb] * 4, dtype=np.int32)
    return Atoms(frac * a, np.eye(3) * a, (True, True, True), Z)


def repeat(at: Atoms, n: Sequence[int]) -> Atoms:
    """`at * (n1,n2,n3)`."""
    if np.isscalar(n):
        n = (int(n),) * 3
    shifts = np.array([[i, j, k] for i in range(n[0]) for j in range(n[1]) for k in range(n[2])],
                      dtype=np.float64)
    X = (at.X[None, :, :] + (shifts @ at.cell)[:, None, :]).reshape(-1, 3)
    Z = np.tile(at.Z, len(shifts))
    cell = at.cell * np.asarray(n, dtype=np.float64)[:, None]
    return Atoms(X, cell, at.pbc, Z)


def rattle(at: Atoms, amp: float, rng: np.random.Generator, strain: float = 0.0, mode: str = "box") -> Atoms:
    """mode "box": every coordinate moves by up to +-amp (the frozen recipe of the fixtures); mode "julip": the
    displacement of JuLIP's `rattle!(at, amp)`, amp * 2/sqrt(3) * (rand(3) - 0.5): at most amp in norm."""
    if mode == "julip":
        X = at.X + amp * (2.0 / np.sqrt(3.0)) * (rng.uniform(0.0, 1.0, size=at.X.shape) - 0.5)
    else:
        X = at.X + amp * rng.uniform(-1.0, 1.0, size=at.X.shape)
    cell = at.cell.copy()
    if strain > 0.0:
        e = rng.uniform(-strain, strain, size=(3, 3))
        Fm = np.eye(3) + 0.5 * (e + e.T)
        X = X @ Fm
        cell = cell @ Fm
    return Atoms(X, cell, at.pbc, at.Z.copy())


# ----------------------------------------------------------------------------
# neighbour enumeration for the label calculators (numpy, all images)
# ----------------------------------------------------------------------------
def _pairs(at: Atoms, rcut: float):
    C = at.cell
    Ci = np.linalg.inv(C)
    rng_ = []
    for a in range(3):
        if not at.pbc[a]:
            rng_.append(np.array([0]))
            continue
        h = 1.0 / np.linalg.norm(Ci[:, a])
        s = at.X @ Ci[:, a]
        n = int(math.floor(rcut / h + (s.max() - s.min())))
        rng_.append(np.arange(-n, n + 1))
    S = np.array([[i, j, k] for i in rng_[0] for j in rng_[1] for k in rng_[2]], dtype=np.float64)
    sh = S @ C                                                # [ns,3]
    D = at.X[None, :, None, :] - at.X[:, None, None, :] + sh[None, None, :, :]   # [i,j,s,3]
    r2 = np.einsum("ijsd,ijsd->ijs", D, D)
    mask = r2 < rcut * rcut
    zero = np.all(S == 0, axis=1)
    ii = np.arange(len(at))
    mask[ii, ii, np.nonzero(zero)[0][0]] = False
    I, J, K = np.nonzero(mask)
    R = D[I, J, K]
    return I, J, R


def _min_distance(at: Atoms) -> float:
    """Smallest interatomic distance (all periodic images)."""
    X, C = at.X, at.cell
    S = np.array([[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)], dtype=np.float64)
    S = S[[all((s == 0) or p for s, p in zip(row, at.pbc)) for row in S]]
    D = X[None, :, None, :] - X[:, None, None, :] + (S @ C)[None, None, :, :]
    r2 = np.einsum("ijsd,ijsd->ijs", D, D)
    ii = np.arange(len(X))
    r2[ii, ii, int(np.nonzero(np.all(S == 0, axis=1))[0][0])] = np.inf
    return float(np.sqrt(r2.min()))


class PairPotential:
    """V_i = 1/2 sum_j phi(r_ij); LennardJones(r0,e0) * C2Shift(rcut) by default
    (`test/test_fit2b.jl:29-31`)."""

    def __init__(self, r0: float, e0: float = 1.0, rcut: Optional[float] = None):
        self.r0, self.e0 = r0, e0
        self.rcut = 2.5 * r0 if rcut is None else rcut
        rc = self.rcut
        self._p = self._lj(rc)

    def _lj(self, r):
        s6 = (self.r0 / r) ** 6
        phi = self.e0 * (s6 * s6 - 2.0 * s6)
        d1 = self.e0 * (-12.0 * s6 * s6 + 12.0 * s6) / r
        d2 = self.e0 * (156.0 * s6 * s6 - 84.0 * s6) / (r * r)
        return phi, d1, d2

    def phi(self, r):
        r = np.asarray(r, dtype=np.float64)
        p, d1, _ = self._lj(r)
        pc, d1c, d2c = self._p
        dr = r - self.rcut
        val = p - pc - d1c * dr - 0.5 * d2c * dr * dr
        dval = d1 - d1c - d2c * dr
        inside = r < self.rcut
        return np.where(inside, val, 0.0), np.where(inside, dval, 0.0)

    def efv(self, at: Atoms):
        I, J, R = _pairs(at, self.rcut)
        r = np.linalg.norm(R, axis=1)
        p, dp = self.phi(r)
        E = 0.5 * p.sum()
        dV = (0.5 * dp / r)[:, None] * R
        F = np.zeros((len(at), 3))
        np.add.at(F, J, -dV)
        np.add.at(F, I, dV)
        V = -np.einsum("pa,pb->ab", dV, R)
        return E, F, V


class StillingerWeber:
    """Stillinger-Weber Si (`test/test_fit3b.jl:29`)."""

    def __init__(self, eps=2.1683, sig=2.0951, a=1.80, lam=21.0, gam=1.20,
                 A=7.049556277, B=0.6022245584, p=4):
        self.eps, self.sig, self.a, self.lam, self.gam, self.A, self.B, self.p = \
            eps, sig, a, lam, gam, A, B, p
        self.rcut = a * sig

    def efv(self, at: Atoms):
        nat = len(at)
        I, J, R = _pairs(at, self.rcut)
        r = np.linalg.norm(R, axis=1)
        Rh = R / r[:, None]
        e, s, rc = self.eps, self.sig, self.rcut
        # two-body
        sr = s / r
        ex = np.exp(s / (r - rc))
        poly = self.B * sr ** self.p - 1.0
        phi2 = self.A * e * poly * ex
        dphi2 = self.A * e * ex * (-self.p * self.B * sr ** self.p / r - poly * s / (r - rc) ** 2)
        E = 0.5 * phi2.sum()
        dV = (0.5 * dphi2)[:, None] * Rh
        # three-body: per site padded lists
        g = np.exp(self.gam * s / (r - rc))
        dg = -g * self.gam * s / (r - rc) ** 2
        order = np.argsort(I, kind="stable")
        I, J, R, r, Rh, g, dg, dV = I[order], J[order], R[order], r[order], Rh[order], g[order], dg[order], dV[order]
        starts = np.searchsorted(I, np.arange(nat + 1))
        for i in range(nat):
            a0, a1 = starts[i], starts[i + 1]
            n = a1 - a0
            if n < 2:
                continue
            rh, gi, dgi, ri = Rh[a0:a1], g[a0:a1], dg[a0:a1], r[a0:a1]
            c = rh @ rh.T
            q = c + 1.0 / 3.0
            off = 1.0 - np.eye(n)
            h = self.lam * e * q * q * np.outer(gi, gi) * off
            E += 0.5 * h.sum()
            # d/dR_j of sum_{k != j} h(j,k)
            ang = 2.0 * self.lam * e * q * np.outer(gi, gi) * off        # [j,k]
            t1 = (ang[:, :, None] * (rh[None, :, :] - c[:, :, None] * rh[:, None, :])).sum(1) / ri[:, None]
            rad = (self.lam * e * q * q * np.outer(dgi, gi) * off).sum(1)
            dV[a0:a1] += t1 + rad[:, None] * rh
        F = np.zeros((nat, 3))
        np.add.at(F, J, -dV)
        np.add.at(F, I, dV)
        V = -np.einsum("pa,pb->ab", dV, R)
        return E, F, V


def rattled_configs(species: str, L: int, nconfigs: int, *, rmax: Optional[float] = None,
                    strain: float = 0.02, seed: int = 1, calc=None, configtype: str = "rand",
                    virial: bool = True, vacancies: int = 0, min_dist: float = 0.0, mode: str = "box") -> List[Dat]:
    """`generate_data(species, L, rmax, N, calc)` of the reference tests, with the
    frozen recipe above.  `vacancies`: remove up to this many random atoms.
    `min_dist` (in units of r0): a rattled cell with a pair closer than this is drawn again (same
    stream), so that a handful of near-collisions cannot dominate the labels."""
    from .basis import rnn
    r0 = rnn(species)
    rmax = 0.33 * r0 if rmax is None else rmax
    rng = np.random.Generator(np.random.Philox(key=seed))
    base = repeat(bulk(species), L)
    out = []
    for _ in range(nconfigs):
        at = rattle(base, rng.uniform() * rmax, rng, strain, mode)
        while min_dist > 0.0 and _min_distance(at) < min_dist * r0:
            at = rattle(base, rng.uniform() * rmax, rng, strain, mode)
        ct = configtype
        if vacancies:
            nv = int(rng.integers(0, vacancies + 1))
            if nv:
                keep = np.ones(len(at), dtype=bool)
                keep[rng.choice(len(at), size=nv, replace=False)] = False
                at = Atoms(at.X[keep], at.cell, at.pbc, at.Z[keep])
                ct = "vacancy"
        if calc is None:
            out.append(Dat(at, ct))
        else:
            E, F, V = calc.efv(at)
            out.append(Dat(at, ct, E=E, F=F, V=V if virial else None))
    return out
