"""Force preconditioners.

Mirrors `/root/reference/src/preconditioners.jl`: `AdjacancyPrecon(intervals; stab, innerstab)`
(`:27-37`), the bond block `k ((1-c) Rhat Rhat' + c I)` for `r0 <= r <= r1` (`:42-52`) and the
assembly over the neighbour list with `j < i` skipped and `P += stab * I` (`:54-74`).
The neighbour list comes from the CUDA K1 kernel (`ipf_neighbours`); the dense 3N x 3N block is
assembled on the host (it is a user-replaceable callable, `weights["precon"][configtype](at)`,
`src/lsq.jl:200-204`); its factorisation and application to the force rows run on the GPU
(`ipf_solve_precon`).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

from . import _capi


class AdjacancyPrecon:
    def __init__(self, intervals: Sequence[Tuple[float, float, float]], stab: float = 1e-4, innerstab: float = 0.1,
                 ctx: _capi.Context = None):
        self.intervals = [(float(a), float(b), float(k)) for (a, b, k) in intervals]
        self.stab = float(stab)
        self.innerstab = float(innerstab)
        self.ctx = ctx

    def cutoff(self) -> float:
        return max(i[1] for i in self.intervals)

    def bond(self, R) -> np.ndarray:
        """3x3 block of one bond (`src/preconditioners.jl:42-52`)."""
        R = np.asarray(R, dtype=np.float64)
        r = np.linalg.norm(R)
        Rh = R / r
        Pi = (1.0 - self.innerstab) * np.outer(Rh, Rh) + self.innerstab * np.eye(3)
        for (r0, r1, k) in self.intervals:
            if r0 <= r <= r1:
                return k * Pi
        return 0.0 * Pi

    def __call__(self, at) -> np.ndarray:
        ctx = self.ctx or _capi.default_context()
        n = len(at)
        I, J, S, R = _capi.neighbours(ctx, at.X, at.cell, at.pbc, self.cutoff())
        keep = J >= I                                   # `if j < i; continue` (`:60`)
        I, J, R = I[keep], J[keep], R[keep]
        r = np.linalg.norm(R, axis=1)
        k = np.zeros(len(r))
        done = np.zeros(len(r), dtype=bool)
        for (r0, r1, kk) in self.intervals:            # first matching interval wins (`:46-50`)
            m = (~done) & (r >= r0) & (r <= r1)
            k[m] = kk
            done |= m
        Rh = R / r[:, None]
        blocks = k[:, None, None] * ((1.0 - self.innerstab) * Rh[:, :, None] * Rh[:, None, :]
                                     + self.innerstab * np.eye(3)[None, :, :])
        P = np.zeros((n, 3, n, 3))
        np.add.at(P, (I, slice(None), J, slice(None)), -blocks)                       # P[i_a, j_b] -= Pij[a,b]
        np.add.at(P, (J, slice(None), I, slice(None)), -blocks.transpose(0, 2, 1))    # P[j_b, i_a] -= Pij[a,b]
        np.add.at(P, (I, slice(None), I, slice(None)), blocks)
        np.add.at(P, (J, slice(None), J, slice(None)), blocks)
        P = P.reshape(3 * n, 3 * n)
        P[np.diag_indices(3 * n)] += self.stab
        return P
