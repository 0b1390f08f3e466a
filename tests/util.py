"""Shared helpers for the parity tests."""
import numpy as np

import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import synth

PSI_RTOL = 1e-10     # north_star: design-matrix entries agree to a relative error of 1e-10


def si_desc(rcut_factor=2.5, width=1.0):
    r0 = ipf.rnn("Si")
    rcut = rcut_factor * r0
    return ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(rcut - width, rcut))


def rows_for(configs):
    """1-based first rows (0 = absent) per observation type, from Dat.rows."""
    out = {}
    for k in ("E", "F", "V"):
        out[k] = np.array([d.rows[k][0] + 1 if k in d.rows else 0 for d in configs], dtype=np.int64)
    return out


def psi_errors(Psi, Ref, configs):
    """Per (column, observation type) max |diff| relative to the max |ref| of that group.
    Forces of a near-perfect crystal cancel to ~0 entry-wise, so the 1e-10 bar is applied
    against the scale of the (column, okey) block, which is what bounds the rounding error
    of any summation order."""
    worst = 0.0
    for k in ("E", "F", "V"):
        idx = np.concatenate([np.arange(d.rows[k][0], d.rows[k][0] + d.rows[k][1])
                              for d in configs if k in d.rows] or [np.zeros(0, dtype=np.int64)])
        if idx.size == 0:
            continue
        a, b = Psi[idx, :], Ref[idx, :]
        scale = np.abs(b).max(axis=0)
        err = np.abs(a - b).max(axis=0)
        rel = np.where(scale > 0, err / np.where(scale > 0, scale, 1.0), err)
        worst = max(worst, float(rel.max()))
    return worst


def psi_entrywise(Psi, Ref, floor=1e-6):
    """Entry-wise relative error |a - b| / |b| next to the block-scaled figure of `psi_errors`, over the entries that are
    not cancellation zeros (|ref| > floor x the largest entry of their column).  Returns (max, 99.9 % quantile, share of
    ALL entries within 1e-10 entry-wise or below 1e-10 x column scale in absolute terms)."""
    scale = np.abs(Ref).max(axis=0)
    scale = np.where(scale > 0, scale, 1.0)
    big = np.abs(Ref) > floor * scale[None, :]
    rel = np.abs(Psi - Ref)[big] / np.abs(Ref)[big]
    diff = np.abs(Psi - Ref)
    ok = (diff <= 1e-10 * np.abs(Ref)) | (diff <= 1e-10 * scale[None, :])
    return float(rel.max()) if rel.size else 0.0, float(np.quantile(rel, 0.999)) if rel.size else 0.0, float(ok.mean())
