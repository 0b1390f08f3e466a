"""Bond-angle permutation-invariant polynomial (PIP) basis specification.

Host-side mirror of the basis constructors the reference README uses
(`/root/reference/README.md:24-50`):

    r0 = rnn(:Si); rcut = 2.5 * r0
    desc = BondAngleDesc("exp(- (r/$r0 - 1.0))", CosCut(rcut-1, rcut))
    B = [ nbpolys(2, D2, 14); nbpolys(3, D3, 11); nbpolys(4, D4, 8) ]

The arithmetic of `nbpolys` lives in NBodyIPs.jl, which is NOT vendored in the
reference tree.  The maths below is therefore builder-defined
and frozen here; the CPU oracle (`oracle/ipf_oracle.c`) and the CUDA kernels
(`csrc/`) both consume the *tables* this module emits, never re-derive them.

Definition.  An N-body basis function with centre atom i acts on every
unordered set J = {j_1,...,j_{N-1}} of distinct neighbour-list entries of i:

    V_b(i) = sum_J  prod_a fcut(r_a) * P_b(x_1..x_{N-1}; w_12, w_13, ...)

with x_a = transform(r_a), w_ac = Rhat_a . Rhat_c, and P_b the S_{N-1}-orbit
sum of one monomial in (x, w) of total degree 1..deg.  The variables are
ordered (x_1..x_{N-1}, w_12, w_13, .., w_23, ..) -- pairs in lexicographic
order.  Basis functions are ordered by (total degree, representative exponent
tuple) so that a lower-degree basis is a prefix of a higher-degree one.
"""
from __future__ import annotations

import itertools
import math
import re
from dataclasses import dataclass, field
from functools import lru_cache
from typing import List, Sequence, Tuple

import numpy as np

TRANSFORM_MORSE = 0      # x = exp(-a (r/r0 - 1))
TRANSFORM_INVPOW = 1     # x = (r0/r)^a

MAX_BODYORDER = 5
MAX_VARS = 10            # (N-1) + (N-1)(N-2)/2 for N = 5
MAX_DEGREE = 20


@dataclass(frozen=True)
class CosCut:
    """fcut(r) = 1 (r<=rc1), 0.5(1+cos(pi (r-rc1)/(rc2-rc1))) (rc1<r<rc2), 0 (r>=rc2).

    Same constructor as the README's `CosCut(rcut-1, rcut)` (`README.md:36`)."""
    rc1: float
    rc2: float

    def __post_init__(self):
        if not (0.0 < self.rc1 < self.rc2):
            raise ValueError("CosCut needs 0 < rc1 < rc2")

    def __call__(self, r):
        r = np.asarray(r, dtype=np.float64)
        s = (r - self.rc1) / (self.rc2 - self.rc1)
        f = 0.5 * (1.0 + np.cos(np.pi * s))
        return np.where(r <= self.rc1, 1.0, np.where(r >= self.rc2, 0.0, f))


_MORSE_RE = re.compile(
    r"^\s*exp\(\s*-\s*(?:(?P<a>[0-9.eE+-]+)\s*\*\s*)?\(\s*r\s*/\s*(?P<r0>[0-9.eE+-]+)\s*-\s*1(?:\.0*)?\s*\)\s*\)\s*$")
_INVPOW_RE = re.compile(r"^\s*\(\s*(?P<r0>[0-9.eE+-]+)\s*/\s*r\s*\)\s*\^\s*(?P<a>[0-9.eE+-]+)\s*$")


@dataclass(frozen=True)
class BondAngleDesc:
    """Bond-angle descriptor: space transform + cutoff (`README.md:34-37`).

    `transform` may be the README's string form ("exp(- (r/2.35 - 1.0))",
    optionally "exp(- 2.0 * (r/2.35 - 1.0))", or "(2.35/r)^3") or a tuple
    (kind, a, r0)."""
    transform: object = field(compare=False)
    cutoff: CosCut = None
    kind: int = field(init=False, default=TRANSFORM_MORSE)
    a: float = field(init=False, default=1.0)
    r0: float = field(init=False, default=1.0)

    def __post_init__(self):
        t = self.transform
        if isinstance(t, str):
            m = _MORSE_RE.match(t)
            if m:
                kind, a, r0 = TRANSFORM_MORSE, float(m.group("a") or 1.0), float(m.group("r0"))
            else:
                m = _INVPOW_RE.match(t)
                if not m:
                    raise ValueError(f"unsupported space transform: {t!r}")
                kind, a, r0 = TRANSFORM_INVPOW, float(m.group("a")), float(m.group("r0"))
        else:
            kind, a, r0 = int(t[0]), float(t[1]), float(t[2])
        if kind not in (TRANSFORM_MORSE, TRANSFORM_INVPOW):
            raise ValueError("unknown transform kind")
        object.__setattr__(self, "kind", kind)
        object.__setattr__(self, "a", a)
        object.__setattr__(self, "r0", r0)

    @property
    def rcut(self) -> float:
        return self.cutoff.rc2

    def x(self, r):
        r = np.asarray(r, dtype=np.float64)
        if self.kind == TRANSFORM_MORSE:
            return np.exp(-self.a * (r / self.r0 - 1.0))
        return (self.r0 / r) ** self.a

    def params(self) -> Tuple[int, float, float, float, float]:
        return (self.kind, self.a, self.r0, self.cutoff.rc1, self.cutoff.rc2)


def nvars(N: int) -> int:
    nx = N - 1
    return nx + nx * (nx - 1) // 2


def _pairs(nx: int):
    return list(itertools.combinations(range(nx), 2))


@lru_cache(maxsize=None)
def invariant_orbits(N: int, deg: int, groups: Tuple[int, ...] = None) -> Tuple[Tuple[Tuple[int, ...], ...], ...]:
    """All S_{N-1}-orbits of monomials of total degree 1..deg in the nvars(N)
    bond-angle variables.  Returns a tuple of orbits, each a sorted tuple of
    exponent tuples; orbits are sorted by (degree, smallest member).
    `groups` (one label per neighbour slot, e.g. the sorted species of the neighbours): only the permutations
    that keep every label in place act, i.e. the polynomial is invariant under exchanging legs of the SAME species."""
    if not (2 <= N <= MAX_BODYORDER):
        raise ValueError("body-order must be 2..5")
    if not (1 <= deg <= MAX_DEGREE):
        raise ValueError("degree must be 1..20")
    nx = N - 1
    pairs = _pairs(nx)
    pidx = {p: k for k, p in enumerate(pairs)}
    nv = nx + len(pairs)
    perms = list(itertools.permutations(range(nx)))
    if groups is not None:
        if len(groups) != nx:
            raise ValueError("one group label per neighbour slot")
        perms = [p for p in perms if all(groups[p[i]] == groups[i] for i in range(nx))]
    # for each permutation, the induced permutation of the variable slots
    slot_maps = []
    for p in perms:
        sm = [p[i] for i in range(nx)]
        for (u, v) in pairs:
            a, b = sorted((p[u], p[v]))
            sm.append(nx + pidx[(a, b)])
        slot_maps.append(sm)

    def gen(k, d):
        if k == 0:
            yield ()
            return
        for e in range(d + 1):
            for rest in gen(k - 1, d - e):
                yield (e,) + rest

    seen = set()
    orbits = []
    for m in gen(nv, deg):
        if sum(m) == 0 or m in seen:
            continue
        orb = {tuple(m[s] for s in sm) for sm in slot_maps}
        seen |= orb
        orbits.append(tuple(sorted(orb)))
    orbits.sort(key=lambda o: (sum(o[0]), o[0]))
    return tuple(orbits)


@dataclass(frozen=True)
class NBPoly:
    """One basis function: body-order N, descriptor, and its monomial orbit."""
    N: int
    desc: BondAngleDesc
    orbit: Tuple[Tuple[int, ...], ...]
    # species-resolved function: centre species zc and the SORTED species of the N-1 neighbours; the function acts on
    # the clusters with exactly these species (legs assigned to the slots in species order) and is zero elsewhere.
    # zc = 0: species-agnostic (every cluster, full S_{N-1} symmetry).
    zc: int = 0
    zs: Tuple[int, ...] = ()
    # environment-dependent function (NBodyIPs `EnvIP`, the `env` of the README's `W5Bdeg12env`, `README.md:65`): the
    # site function is multiplied by n_i^env_t, n_i = sum_j fenv(r_ij) the smooth coordination number of the centre
    # atom with fenv = CosCut(env_rc[0], env_rc[1]) (env_rc[1] <= the descriptor's cutoff).  env_t = 0: plain function.
    env_t: int = 0
    env_rc: Tuple[float, float] = (0.0, 0.0)

    @property
    def degree(self) -> int:
        return sum(self.orbit[0])


def bodyorder(b: NBPoly) -> int:
    return b.N


def degree(b: NBPoly) -> int:
    return b.degree


def nbpolys(N: int, desc: BondAngleDesc, deg: int, species=None) -> List[NBPoly]:
    """`nbpolys(bodyorder, descriptor, degree)` (`README.md:38-42`).
    `species = (zc, (z_1, .., z_{N-1}))`: the functions of ONE species tuple (centre, neighbours); the polynomials are
    invariant under permutations of equal-species neighbours only.  `species_nbpolys` builds all tuples."""
    if species is None:
        return [NBPoly(N, desc, orb) for orb in invariant_orbits(N, deg)]
    zc, zs = int(species[0]), tuple(sorted(int(z) for z in species[1]))
    if zc <= 0 or len(zs) != N - 1 or min(zs) <= 0:
        raise ValueError("species = (zc, (z_1, .., z_{N-1})) with positive atomic numbers")
    return [NBPoly(N, desc, orb, zc, zs) for orb in invariant_orbits(N, deg, zs)]


def envpolys(N: int, desc: BondAngleDesc, deg: int, tmax: int, envcut: CosCut, species=None) -> List[NBPoly]:
    """Environment-dependent N-body basis: for t = 0..tmax the functions n_i^t * nbpolys(N, desc, deg) (t-major), with the
    smooth coordination number n_i = sum_j envcut(r_ij).  Builder-defined restatement of NBodyIPs' `envpolys` (the maths is
    not vendored in the reference; `README.md:65` only names a 5-body `env` database)."""
    if tmax < 0 or not (envcut.rc2 <= desc.rcut):
        raise ValueError("tmax >= 0 and the environment cutoff must not exceed the descriptor cutoff")
    base = nbpolys(N, desc, deg, species)
    out: List[NBPoly] = []
    for t in range(tmax + 1):
        out += [NBPoly(p.N, p.desc, p.orbit, p.zc, p.zs, t, (float(envcut.rc1), float(envcut.rc2)) if t > 0 else (0.0, 0.0))
                for p in base]
    return out


def species_nbpolys(N: int, desc: BondAngleDesc, deg: int, zlist: Sequence[int]) -> List[NBPoly]:
    """The species-resolved N-body basis of a multi-component system (JuLIP `Atoms.Z`, `src/prototypes.jl:19-25`): one
    run of functions per (centre species, multiset of neighbour species), centre-major, multisets in lexicographic order."""
    out: List[NBPoly] = []
    zl = sorted(int(z) for z in zlist)
    for zc in zl:
        for zs in itertools.combinations_with_replacement(zl, N - 1):
            out += nbpolys(N, desc, deg, (zc, zs))
    return out


@dataclass
class BasisBlock:
    """A run of consecutive basis functions sharing (N, desc): the unit the
    C-ABI and the kernels work on."""
    N: int
    desc: BondAngleDesc
    col0: int                  # first column (0-based) in the full basis
    nb: int                    # number of basis functions
    mono_b: np.ndarray         # int32[nmono]   local basis index of each monomial
    mono_exp: np.ndarray       # uint8[nmono, MAX_VARS] exponents (zero padded)
    maxdeg: int
    zc: int = 0                # centre species (0 = any)
    zs: Tuple[int, ...] = ()   # sorted neighbour species (empty = any)
    env_t: int = 0             # power of the smooth coordination number (0 = plain function)
    env_rc: Tuple[float, float] = (0.0, 0.0)


class NBodyBasis:
    """An `IPBasis`: ordered list of NBPoly, compiled into per-(N,desc) blocks.

    `len(basis)` is the number of columns of the LSQ matrix
    (`src/lsq_db.jl:249`)."""

    def __init__(self, polys: Sequence[NBPoly]):
        self.polys: List[NBPoly] = list(polys)
        if not self.polys:
            raise ValueError("empty basis")
        self.blocks: List[BasisBlock] = []
        start = 0
        while start < len(self.polys):
            p0 = self.polys[start]
            end = start
            while (end < len(self.polys) and self.polys[end].N == p0.N
                   and self.polys[end].desc == p0.desc and self.polys[end].zc == p0.zc
                   and self.polys[end].zs == p0.zs and self.polys[end].env_t == p0.env_t
                   and self.polys[end].env_rc == p0.env_rc):
                end += 1
            self.blocks.append(self._compile(start, end))
            start = end

    def _compile(self, start: int, end: int) -> BasisBlock:
        p0 = self.polys[start]
        nv = nvars(p0.N)
        mb, me = [], []
        maxdeg = 0
        for k in range(start, end):
            for m in self.polys[k].orbit:
                if len(m) != nv:
                    raise ValueError("monomial has wrong number of variables")
                mb.append(k - start)
                row = list(m) + [0] * (MAX_VARS - nv)
                me.append(row)
                maxdeg = max(maxdeg, sum(m))
        return BasisBlock(N=p0.N, desc=p0.desc, col0=start, nb=end - start,
                          mono_b=np.asarray(mb, dtype=np.int32),
                          mono_exp=np.asarray(me, dtype=np.uint8).reshape(-1, MAX_VARS),
                          maxdeg=maxdeg, zc=p0.zc, zs=tuple(p0.zs), env_t=p0.env_t, env_rc=tuple(p0.env_rc))

    def __len__(self) -> int:
        return len(self.polys)

    def __getitem__(self, i):
        return self.polys[i]

    def __iter__(self):
        return iter(self.polys)

    @property
    def rcut(self) -> float:
        return max(b.desc.rcut for b in self.blocks)

    def __eq__(self, other):
        return isinstance(other, NBodyBasis) and self.polys == other.polys

    # ---- (de)serialisation, same role as JuLIP FIO write_dict/read_dict used
    # ---- by `src/lsq_db.jl:96-104`
    def write_dict(self) -> dict:
        descs, index = [], {}
        polys = []
        for p in self.polys:
            if p.desc not in index:
                index[p.desc] = len(descs)
                descs.append({"kind": p.desc.kind, "a": p.desc.a, "r0": p.desc.r0,
                              "rc1": p.desc.cutoff.rc1, "rc2": p.desc.cutoff.rc2})
            e = {"N": p.N, "desc": index[p.desc], "orbit": [list(m) for m in p.orbit]}
            if p.zc:
                e["zc"] = p.zc
                e["zs"] = list(p.zs)
            if p.env_t:
                e["env_t"] = p.env_t
                e["env_rc"] = list(p.env_rc)
            polys.append(e)
        return {"__id__": "IPFittingB200.NBodyBasis", "descs": descs, "polys": polys}

    @staticmethod
    def read_dict(D: dict) -> "NBodyBasis":
        descs = [BondAngleDesc((d["kind"], d["a"], d["r0"]), CosCut(d["rc1"], d["rc2"]))
                 for d in D["descs"]]
        polys = [NBPoly(p["N"], descs[p["desc"]], tuple(tuple(m) for m in p["orbit"]), int(p.get("zc", 0)),
                        tuple(int(z) for z in p.get("zs", ())), int(p.get("env_t", 0)),
                        tuple(float(v) for v in p.get("env_rc", (0.0, 0.0))))
                 for p in D["polys"]]
        return NBodyBasis(polys)


def as_basis(B) -> NBodyBasis:
    if isinstance(B, NBodyBasis):
        return B
    return NBodyBasis(list(B))


def rnn(species: str) -> float:
    """Nearest-neighbour distance of the ground-state crystal (JuLIP `rnn`).
    Only the species the BASELINE configs use."""
    table = {"Si": math.sqrt(3.0) / 4.0 * 5.43, "Cu": 3.61 / math.sqrt(2.0),
             "W": math.sqrt(3.0) / 2.0 * 3.165, "C": math.sqrt(3.0) / 4.0 * 3.567}
    return table[species]
