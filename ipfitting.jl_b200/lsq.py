"""`lsqfit` and friends: observations + weights -> weighted system -> solve -> IP, fitinfo.

Mirrors `/root/reference/src/lsq.jl`:
  collect_observations   :74-109     _get_weights :118-161     _fix_weights! :698-707
  get_lsq_system         :245-311    lsqfit       :404-614     asm_fitinfo   :617-678
The dense work (row weighting + column scaling `:306-307,450-458`, the QR solve
`:466-473`, `c = D_inv*c` `:593`) runs on the GPU through the C-ABI
(`ipf_solve_*`); everything here is the host-side dictionary logic of the reference.

Only the `:qr` solver family is provided by the CUDA path (`:rrqr`,
`:svd` are "next"; `:lsqr/:brr/:ard/:xgboost` wrap third-party python/Julia solvers
and are out of scope).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import numpy as np

from . import _capi, parallel
from .datatypes import weighthook
from .lsq_db import LsqDB, matrows
from .obsiter import observations
from .potentials import linear_part, NBodyIP, OneBody, SumIP, vref_observation
from .prototypes import ENERGY, FORCES, VIRIAL

_QR = ("qr", ":qr")


def _fix_weights(weights: Optional[dict]) -> dict:
    """`_fix_weights!` (`src/lsq.jl:698-707`)."""
    weights = {} if weights is None else weights
    if "ignore" not in weights:
        weights = dict(weights)
        weights["ignore"] = []
    if "default" not in weights:
        weights["default"] = {"E": 1.0, "F": 1.0, "V": 1.0}
    return weights


def _get_weights(weights: dict, wh: float, dat, obskey: str, o: np.ndarray):
    """`_get_weights` (`src/lsq.jl:118-161`).  Returns a scalar (broadcast over the block) or,
    for E / V with a per-component `info["W"]` entry, a vector of length len(o)."""
    w = 0.0
    if "W" in dat.info:
        # overrides everything, weighthook included (`:125-131`)
        if obskey in dat.info["W"]:
            w = dat.info["W"][obskey]
    else:
        ct = dat.configtype
        cfgkey = ct if ct in weights else "default"
        if obskey in weights[cfgkey]:
            w = weights[cfgkey][obskey] * wh
    if isinstance(w, (int, float)):
        return float(w)
    wa = np.atleast_1d(np.asarray(w, dtype=np.float64))
    if wa.size == 1:
        return float(wa[0])
    if obskey in (ENERGY, VIRIAL) and wa.size == len(o):
        return wa.copy()
    if obskey == FORCES:
        raise ValueError("_get_weights: a vector weight is not allowed for \"F\"")
    raise ValueError("_get_weights: length(w) is neither 1 nor length(o)?!?!?")


_OKEYS = (ENERGY, FORCES, VIRIAL)


def _observation_layout(db):
    """Per-observation arrays in `observations(db)` order, cached on the db (rows never change after
    `_alloc_lsq_matrix`): config index, okey id, first row, length, natoms, configtype id."""
    lay = getattr(db, "_obs_layout", None)
    if lay and lay.get("nconfigs") == len(db.configs):
        return lay
    cfg, key, first, length, nat, ctid = [], [], [], [], [], []
    cts = {}
    plain = True
    for icfg, d in enumerate(db.configs):
        ci = cts.setdefault(d.configtype, len(cts))
        n = len(d.at)
        for okey in d.D.keys():
            if okey not in _OKEYS:
                plain = False
                continue
            f, ln = d.rows[okey]
            cfg.append(icfg); key.append(_OKEYS.index(okey)); first.append(f); length.append(ln)
            nat.append(n); ctid.append(ci)
    lay = {"nconfigs": len(db.configs), "plain": plain, "cts": list(cts.keys()),
           "cfg": np.asarray(cfg, dtype=np.int64), "key": np.asarray(key, dtype=np.int64),
           "first": np.asarray(first, dtype=np.int64), "len": np.asarray(length, dtype=np.int64),
           "nat": np.asarray(nat, dtype=np.float64), "ctid": np.asarray(ctid, dtype=np.int64)}
    # rows in traversal order are contiguous and ascending (that is how they were assigned)
    lay["contiguous"] = bool(len(first) == 0 or (np.array_equal(lay["first"][1:], lay["first"][:-1] + lay["len"][:-1])
                                                  and lay["first"][0] == 0))
    try:
        db._obs_layout = lay
    except Exception:
        pass
    return lay


def collect_observations(db: LsqDB, weights: dict, Vref):
    """`collect_observations(db, weights, Vref) -> Y, W, Icfg` (`src/lsq.jl:74-109`).
    Icfg is 0-based here, -1 for rows that were skipped.  Vectorised when no configuration carries
    `info["W"]` / `info["Vref"]` overrides and Vref is None or a OneBody; otherwise the
    observation-by-observation loop of the reference."""
    nrows = db.nrows
    ignore = weights["ignore"]
    onebody = isinstance(Vref, OneBody)
    lay = _observation_layout(db)
    # a reference potential that is linear in one of OUR bases (a previous fit, possibly + OneBody terms) is evaluated
    # for all configurations at once: assemble its design matrix on the device, multiply by its coefficients
    lin, others = (None, []) if (Vref is None or onebody) else linear_part(Vref)
    batched = lin is not None and all(isinstance(o, OneBody) and not isinstance(o.E0, dict) for o in others)
    simple = (lay["plain"] and lay["contiguous"] and (Vref is None or onebody or batched) and int(lay["len"].sum()) == nrows
              and not any(("W" in d.info) or ("Vref" in d.info) for d in db.configs))
    if simple:
        Y = (np.concatenate([v for d in db.configs for v in d.D.values()]) if nrows else np.zeros(0)).astype(np.float64, copy=False)
        nobs = len(lay["cfg"])
        wobs = np.zeros(nobs)
        live = np.ones(nobs, dtype=bool)
        hook = np.where(lay["key"] == 1, 1.0, 1.0 / np.sqrt(lay["nat"]))       # weighthook: E, V -> 1/sqrt(N)
        for ci, ct in enumerate(lay["cts"]):
            sel_ct = lay["ctid"] == ci
            if ct in ignore:
                live &= ~sel_ct
                continue
            table = weights[ct] if ct in weights else weights["default"]
            for ki, okey in enumerate(_OKEYS):
                if okey in table:
                    m = sel_ct & (lay["key"] == ki)
                    wobs[m] = float(table[okey]) * hook[m]
        for ki, okey in enumerate(_OKEYS):
            if okey in ignore:
                live &= lay["key"] != ki
        if onebody:
            esel = (lay["key"] == 0) & live
            if esel.any():
                if isinstance(Vref.E0, dict):
                    e0 = np.array([Vref.energy(db.configs[i].at) for i in lay["cfg"][esel]])
                else:                           # E0 * natoms, the same product OneBody.energy forms
                    e0 = float(Vref.E0) * lay["nat"][esel]
                Y[lay["first"][esel]] -= e0        # Y is a fresh array (concatenate above)
        if batched and nrows:
            from .lsq_db import LsqDB as _LsqDB
            dbr = _LsqDB("", lin.basis, db.configs, ctx=getattr(db, "ctx", None))      # same configs => same rows
            cx = dbr._context()
            cr = np.ascontiguousarray(lin.c, dtype=np.float64)
            yref = np.empty(dbr.nrows)
            cx.check(cx.lib.ipf_predict(cx.h, dbr.Psi_dev.h, _capi.ptr(cr, C.c_double), _capi.ptr(yref, C.c_double)))
            dbr.delete()
            esel = lay["key"] == 0
            for o in others:
                yref[lay["first"][esel]] += float(o.E0) * lay["nat"][esel]
            Y -= yref
        W = np.repeat(np.where(live, wobs, 0.0), lay["len"])
        Icfg = np.repeat(np.where(live, lay["cfg"], -1), lay["len"])
        if not live.all():
            Y = np.where(np.repeat(live, lay["len"]), Y, 0.0)
        return Y, W, Icfg
    Y = np.zeros(nrows)
    W = np.zeros(nrows)
    Icfg = np.full(nrows, -1, dtype=np.int64)
    for icfg, dat in enumerate(db.configs):
        if dat.configtype in ignore:
            continue
        info = dat.info
        for okey in list(dat.D.keys()):          # same traversal as observations(db) (`src/obsiter.jl:36-59`)
            if okey in ignore:
                continue
            f, n = dat.rows[okey]
            obs = dat.D[okey]
            if "Vref" in info:
                obs = obs - np.asarray(info["Vref"][okey], dtype=np.float64)
            elif onebody:
                if okey == ENERGY:               # OneBody: zero forces and virial
                    obs = obs - Vref.energy(dat.at)
            elif Vref is not None:
                obs = obs - vref_observation(Vref, okey, dat)
            Y[f:f + n] = obs
            Icfg[f:f + n] = icfg
            W[f:f + n] = _get_weights(weights, weighthook(okey, dat), dat, okey, obs)
    return Y, W, Icfg


def _select(db: LsqDB, Y, W, Icfg, Ibasis, Itrain):
    """Row/column selection of `get_lsq_system` (`src/lsq.jl:270-283`)."""
    # a NaN anywhere makes the sum NaN (no temporary); inf - inf can too, hence the exact re-check
    if np.isnan(Y.sum()) and np.isnan(Y).any():
        raise _capi.IPFError(-4, "NaN detected in observations vector")
    if np.isnan(W.sum()) and np.isnan(W).any():
        raise _capi.IPFError(-5, "NaN detected in weights vector")
    if Itrain is None:
        Irows = np.flatnonzero(W)            # NaN != 0 is excluded above
    else:
        keep = W != 0
        keep &= np.isin(Icfg, np.asarray(list(Itrain), dtype=np.int64))
        Irows = np.flatnonzero(keep)
    Irows = Irows.astype(np.int64, copy=False)
    Icols = None if Ibasis is None else np.asarray(list(Ibasis), dtype=np.int64)
    return Irows, Icols


def _precon_blocks(db, weights, Irows):
    """The (row position in the selected system, matrix) pairs `_forceprecon` acts on (`src/lsq.jl:184-234`):
    every "F" observation whose configtype (or "default") has an entry in `weights["precon"]`."""
    pre = weights.get("precon")
    if not pre:
        return [], [], "chol"
    solver = pre.get("_solver", "chol")
    if solver not in ("chol", "sqrt"):
        raise ValueError("unknown root solver")
    row0, blocks = [], []
    for okey, dat, _ in observations(db):
        if okey != FORCES:
            continue
        ct = dat.configtype
        fn = pre.get(ct, pre.get("default"))
        if fn is None:
            continue
        f, n = dat.rows[okey]
        pos = int(np.searchsorted(Irows, f))
        if pos >= len(Irows) or Irows[pos] != f:
            continue                                  # the force rows carry zero weight: not part of the system
        if pos + n > len(Irows) or Irows[pos + n - 1] != f + n - 1:
            raise ValueError("force block only partially selected")
        P = np.asarray(fn(dat.at), dtype=np.float64)
        if P.shape != (n, n):
            raise ValueError("preconditioner has the wrong shape")
        if solver == "sqrt":
            import scipy.linalg as sla
            P = np.real(sla.sqrtm(P))
        row0.append(pos)
        blocks.append(P)
    return row0, blocks, solver


class _Solver:
    """RAII wrapper of the staged `ipf_solve_*` calls."""

    def __init__(self, ctx: _capi.Context, Psi_dev: _capi.DeviceMatrix, Y, W, Irows, Icols, Dinv):
        self.ctx = ctx
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        W = np.ascontiguousarray(W, dtype=np.float64)
        h = C.c_void_p()
        ir = None if Irows is None else np.ascontiguousarray(Irows, dtype=np.int64)
        ic = None if Icols is None else np.ascontiguousarray(Icols, dtype=np.int64)
        dv = None if Dinv is None else np.ascontiguousarray(Dinv, dtype=np.float64)
        ctx.check(ctx.lib.ipf_solve_begin(
            ctx.h, Psi_dev.h, _capi.ptr(Y, C.c_double), _capi.ptr(W, C.c_double),
            None if ir is None else _capi.ptr(ir, C.c_int64), 0 if ir is None else len(ir),
            None if ic is None else _capi.ptr(ic, C.c_int64), 0 if ic is None else len(ic),
            None if dv is None else _capi.ptr(dv, C.c_double), C.byref(h)))
        self.h = h
        M, n = C.c_int64(), C.c_int64()
        ctx.lib.ipf_solve_shape(h, C.byref(M), C.byref(n))
        self.M, self.n = M.value, n.value

    def apply_precon(self, row0, blocks, solver: str = "chol"):
        """`_forceprecon` (`src/lsq.jl:184-234`) on the device: block b (an m x m matrix P_b, or sqrt(P_b) for
        solver "sqrt") acts on the rows [row0[b], row0[b] + m) of the weighted system."""
        if solver not in ("chol", "sqrt"):
            raise ValueError("unknown root solver")
        if not blocks:
            return
        m = np.array([b.shape[0] for b in blocks], dtype=np.int32)
        off = np.zeros(len(blocks), dtype=np.int64)
        np.cumsum(m[:-1].astype(np.int64) ** 2, out=off[1:])
        P = np.concatenate([np.asfortranarray(b, dtype=np.float64).reshape(-1, order="F") for b in blocks])
        r0 = np.ascontiguousarray(row0, dtype=np.int64)
        self.ctx.check(self.ctx.lib.ipf_solve_precon(
            self.ctx.h, self.h, len(blocks), _capi.ptr(r0, C.c_int64), _capi.ptr(m, C.c_int32), _capi.ptr(off, C.c_int64),
            _capi.ptr(P, C.c_double), len(P), 0 if solver == "chol" else 1))

    def append_rows(self, P, q=None):
        """Regulariser rows [P | q] below the weighted system (`_regularise!`, `src/lsq.jl:163-181`)."""
        P = np.asfortranarray(P, dtype=np.float64)
        if P.ndim != 2 or P.shape[1] != self.n:
            raise ValueError("regulariser has the wrong number of columns")
        if P.shape[0] == 0:
            return
        qq = None if q is None else np.ascontiguousarray(q, dtype=np.float64)
        if qq is not None and qq.shape != (P.shape[0],):
            raise ValueError("regulariser target has the wrong length")
        self.ctx.check(self.ctx.lib.ipf_solve_append_rows(
            self.ctx.h, self.h, P.shape[0], _capi.ptr(P, C.c_double), P.shape[0],
            None if qq is None else _capi.ptr(qq, C.c_double)))
        self.M += P.shape[0]

    def system_to_host(self):
        A = np.empty((self.M, self.n), order="F")
        y = np.empty(self.M)
        self.ctx.check(self.ctx.lib.ipf_solve_system_to_host(
            self.ctx.h, self.h, _capi.ptr(A, C.c_double), max(1, self.M), _capi.ptr(y, C.c_double)))
        return A, y

    def gram(self):
        g, n1 = C.c_void_p(), C.c_int64()
        self.ctx.check(self.ctx.lib.ipf_solve_gram(self.ctx.h, self.h, C.byref(g), C.byref(n1)))
        return g.value, n1.value

    def factor(self) -> bool:
        done = C.c_int32()
        self.ctx.check(self.ctx.lib.ipf_solve_factor(self.ctx.h, self.h, C.byref(done)))
        return bool(done.value)

    def finish(self, want_R=True):
        c = np.empty(self.n)
        R = np.empty((self.n, self.n), order="F") if want_R else None
        ssr, ssy = C.c_double(), C.c_double()
        self.ctx.check(self.ctx.lib.ipf_solve_finish(
            self.ctx.h, self.h, _capi.ptr(c, C.c_double), None if R is None else _capi.ptr(R, C.c_double),
            C.byref(ssr), C.byref(ssy)))
        return c, R, ssr.value, ssy.value

    def qty(self) -> np.ndarray:
        """Q^T y of the factorisation A0 = Q R (after the factorisation is complete)."""
        q = np.empty(self.n)
        self.ctx.check(self.ctx.lib.ipf_solve_qty(self.ctx.h, self.h, _capi.ptr(q, C.c_double)))
        return q

    def residual(self, c):
        """(ssr, ssy) = (|y - A0 c|^2, |y|^2) over the local rows, c in the scaled basis (before D_inv)."""
        c = np.ascontiguousarray(c, dtype=np.float64)
        ssr, ssy = C.c_double(), C.c_double()
        self.ctx.check(self.ctx.lib.ipf_solve_residual(self.ctx.h, self.h, _capi.ptr(c, C.c_double),
                                                       C.byref(ssr), C.byref(ssy)))
        return ssr.value, ssy.value

    def set_csne(self, on: bool):
        self.ctx.check(self.ctx.lib.ipf_solve_set_csne(self.h, 1 if on else 0))

    def cond(self) -> float:
        """cond(R) of the returned factor, estimated on the device (power / inverse iteration; after `finish`)."""
        k = C.c_double()
        self.ctx.check(self.ctx.lib.ipf_solve_cond(self.ctx.h, self.h, C.byref(k)))
        return k.value

    def info(self) -> dict:
        npass, defect, shift, gms, tms = C.c_int32(), C.c_double(), C.c_double(), C.c_double(), C.c_double()
        self.ctx.lib.ipf_solve_info(self.h, C.byref(npass), C.byref(defect), C.byref(shift), C.byref(gms), C.byref(tms))
        return {"passes": npass.value, "orth_defect": defect.value, "shift": shift.value,
                "kappa_hat": gms.value, "csne_steps": int(tms.value)}

    def close(self):
        if self.h:
            self.ctx.lib.ipf_solve_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_staged(S, comm=None, device: int = 0, want_R: bool = True, factor_solver=None, Dinv=None):
    """The staged solve protocol (include/ipf_b200.h): gram -> [all-reduce] -> factor, repeated
    until done; then finish.  `S` is a `_Solver` (device) -- or any object with the same four
    methods (the gloo tests drive a numpy stand-in through exactly this function).
    `factor_solver(R, qty) -> (c_scaled, info)` replaces the triangular solve by a host-side solve on the factor."""
    multi = comm is not None and comm.world_size > 1
    if multi and not (hasattr(comm, "attached") and comm.attached()) and hasattr(S, "set_csne"):
        S.set_csne(False)          # the Gram blocks are summed out here: the library only sees local rows in finish
    while True:
        g, n1 = S.gram()
        if multi:
            comm.allreduce_gram(g, n1 * n1, device)
        if S.factor():
            break
    c, R, ssr, ssy = S.finish(want_R or factor_solver is not None)
    extra = {}
    if factor_solver is not None:
        # :svd / :rrqr: the solve happens on the n x n factor (host), the residual on the device
        cs, extra = factor_solver(R, S.qty())
        ssr, ssy = S.residual(cs)
        c = cs if Dinv is None else cs * Dinv
    if multi and not (hasattr(comm, "attached") and comm.attached()):
        ssr, ssy = comm.allreduce_scalars([ssr, ssy])      # (an attached library communicator has summed them already)
    info = S.info()
    info.update(extra)
    info["rel_rms"] = math.sqrt(ssr) / math.sqrt(ssy) if ssy > 0 else float("nan")
    if hasattr(S, "cond"):
        info["cond_R_estimate"] = S.cond()
    return c, R, info


def solve_device(ctx, Psi_dev, Y, W, Irows=None, Icols=None, Dinv=None, want_R=True, comm=None, precon=None,
                 factor_solver=None, reg=None, exchange="gram"):
    """Weighted LSQ solve of the device-resident system.  `comm`: a
    `parallel.Communicator` (rows are sharded over its ranks) or None.  `precon`: (row0, blocks, solver).
    `exchange`: what the ranks exchange -- "gram" (default: the (n+1)^2 Gram block of every pass is summed) or "tsqr"
    (every rank factors its own rows, the blocks [R_p | Q_p^T y] are exchanged once and the stacked system is factored
    again on every rank)."""
    S = _Solver(ctx, Psi_dev, Y, W, Irows, Icols, Dinv)
    try:
        if precon is not None and precon[1]:
            S.apply_precon(*precon)
        if reg is not None and reg[0] is not None and (comm is None or comm.rank == 0):
            S.append_rows(*reg)          # one rank carries the regulariser rows
        if exchange == "tsqr" and comm is not None and comm.world_size > 1:
            if factor_solver is not None:
                raise NotImplementedError("the TSQR exchange is implemented for the :qr solver")
            return _run_tsqr(S, ctx, comm, Dinv)
        return run_staged(S, comm, ctx.device, want_R, factor_solver, Dinv)
    finally:
        S.close()


def _run_tsqr(S, ctx, comm, Dinv):
    """Row-sharded TSQR with CholeskyQR as the local factorisation: A_p = Q_p R_p on every rank (no exchange), then
    min |[R_1; ..; R_N] c - [Q_1^T y; ..; Q_N^T y]| on every rank.  One exchange of N n (n + 1) doubles instead of one
    (n + 1)^2 block per Gram pass; the local factors never see the other ranks' rows, so a rank whose rows alone are rank
    deficient still contributes a valid (singular) R_p."""
    paused = comm.attached(ctx)
    if paused:
        ctx.comm_pause(True)
    try:
        _, Rp, info1 = run_staged(S, None, ctx.device, True, None, None)
        qty = S.qty()
        _, ssy_p = S.residual(np.zeros(S.n))           # |y_p|^2 of the local rows
    finally:
        if paused:
            ctx.comm_pause(False)
    n, N = S.n, comm.world_size
    buf = np.zeros((N, n, n + 1))
    buf[comm.rank, :, :n] = np.triu(Rp)
    buf[comm.rank, :, n] = qty
    buf = comm.allreduce_numpy(buf)
    A2 = np.asfortranarray(buf[:, :, :n].reshape(N * n, n))
    y2 = np.ascontiguousarray(buf[:, :, n].reshape(N * n))
    P2 = _capi.DeviceMatrix(ctx, N * n, n)
    try:
        P2.from_host(A2)
        S2 = _Solver(ctx, P2, y2, np.ones(N * n), None, None, None)
        try:
            if paused:
                ctx.comm_pause(True)           # the second level is the same on every rank: no exchange either
            try:
                c2, R2, info2 = run_staged(S2, None, ctx.device, True, None, None)
                ssr2 = info2["rel_rms"] ** 2 * float(y2 @ y2)
            finally:
                if paused:
                    ctx.comm_pause(False)
        finally:
            S2.close()
    finally:
        P2.close()
    # |A c - y|^2 = |R_stack c - q_stack|^2 + sum_p (|y_p|^2 - |Q_p^T y_p|^2)
    ssy, tail = comm.allreduce_scalars([ssy_p, max(0.0, ssy_p - float(qty @ qty))])
    info = dict(info2)
    info["passes_local"] = info1.get("passes")
    info["exchange"] = "tsqr"
    info["rel_rms"] = math.sqrt((ssr2 + tail) / ssy) if ssy > 0 else float("nan")
    c = c2 if Dinv is None else c2 * Dinv
    return c, R2, info


def get_lsq_system(db: LsqDB, *, verbose=False, Ibasis=None, Itrain=None, E0=None, Vref="default",
                   weights=None, regularisers=()):
    """`get_lsq_system(db; ...) -> Psi_W, Y_W` on the host (`src/lsq.jl:245-311`); default
    `Vref = OneBody(E0)` as there."""
    if isinstance(Vref, str) and Vref == "default":
        Vref = OneBody(E0 if E0 is not None else 0.0)
    weights = _fix_weights(weights)
    Y, W, Icfg = collect_observations(db, weights, Vref)
    Irows, Icols = _select(db, Y, W, Icfg, Ibasis, Itrain)
    S = _Solver(db._context(), db.Psi_dev, Y, W, Irows, Icols, None)
    try:
        row0, blocks, psolver = _precon_blocks(db, weights, Irows)
        S.apply_precon(row0, blocks, psolver)
        if regularisers:
            S.append_rows(*_regulariser_rows(regularisers, db.basis, Icols))
        return S.system_to_host()
    finally:
        S.close()


def _svd_factor_solver(ndiscard: int):
    """`:svd` (`src/lsq.jl:482-488`): c = V[:, 1:end-k] * (S[1:end-k] \\ (U'Y)[1:end-k]).  With Psi = Q R the SVD of
    the n x n factor gives the same S, V and U'Y = U_R' (Q'Y)."""
    def solve(R, qty):
        n = R.shape[0]
        if not 0 <= ndiscard < max(n, 1):
            raise ValueError("svd_ndiscard out of range")
        U, S, Vt = np.linalg.svd(np.triu(R))
        k = n - ndiscard
        c = Vt[:k].T @ ((U.T @ qty)[:k] / S[:k])
        return c, {"svd_S": S}
    return solve


def _rrqr_factor_solver(rtol: float):
    """`:rrqr` (`src/lsq.jl:489-496`, `pqrfact(Psi, rtol=...)` of LowRankApprox.jl): column-pivoted QR truncated
    where |R_kk| <= rtol |R_11|, basic solution (zeros for the discarded columns).  The pivoted QR of the n x n
    factor selects the same columns as that of Psi = Q R (equal column norms at every step)."""
    def solve(R, qty):
        import scipy.linalg as sla
        n = R.shape[0]
        Q2, R2, piv = sla.qr(np.triu(R), pivoting=True)
        d = np.abs(np.diag(R2))
        k = int(np.sum(d > rtol * d[0])) if n else 0
        c = np.zeros(n)
        if k:
            c[piv[:k]] = sla.solve_triangular(R2[:k, :k], (Q2.T @ qty)[:k])
        return c, {"rrqr_rank": k}
    return solve


def _matrix_p_factor_solver(B: np.ndarray):
    """`solver["P"]` a full matrix (`src/lsq.jl:450-458,593`: D_inv = pinv(P); Psi <- Psi D_inv; qr; c = D_inv (qr \\ Y)).
    The device factors the matrix WITHOUT the column transformation, Psi = Q R; then Psi D_inv = Q (R D_inv) and the QR
    factorisation of the n x n product R D_inv = Q2 R2 gives the factor of the reference's matrix, R2, its right-hand
    side Q2' (Q'Y), and c = D_inv R2^-1 Q2' Q'Y -- n x n host work instead of an M x n x n product on Psi."""
    def solve(R, qty):
        import scipy.linalg as sla
        Q2, R2 = np.linalg.qr(np.triu(R) @ B)
        sgn = np.where(np.diag(R2) < 0, -1.0, 1.0)           # a factor with a positive diagonal, as Householder gives up to sign
        R2 = sgn[:, None] * R2
        cp = sla.solve_triangular(R2, sgn * (Q2.T @ qty))
        return B @ cp, {"R_factor": R2}
    return solve


def _regulariser_rows(regularisers, basis, Icols):
    """`_regularise!` (`src/lsq.jl:163-181`): every regulariser is a matrix P (target 0), a pair (P, q), or an object
    with `.matrix(basis) -> (P, q)` (the reference's `Matrix(reg, basis)`); all are stacked.  With `Ibasis` the
    selected columns are taken (the reference leaves that case as a TODO)."""
    Ps, qs = [], []
    for reg in regularisers:
        if hasattr(reg, "matrix") and not isinstance(reg, np.ndarray):
            P, q = reg.matrix(basis)
        elif isinstance(reg, (tuple, list)) and len(reg) == 2 and np.ndim(reg[0]) == 2:
            P, q = reg
        else:
            P, q = reg, None
        P = np.asarray(P, dtype=np.float64)
        if P.ndim != 2 or P.shape[1] != len(basis):
            raise ValueError("regulariser matrix must have one column per basis function")
        q = np.zeros(P.shape[0]) if q is None else np.asarray(q, dtype=np.float64).reshape(-1)
        if q.shape[0] != P.shape[0]:
            raise ValueError("regulariser target has the wrong length")
        Ps.append(P if Icols is None else P[:, Icols]); qs.append(q)
    if not Ps:
        return None, None
    return np.vstack(Ps), np.concatenate(qs)


def _is_full_matrix(P) -> bool:
    P = np.asarray(P, dtype=np.float64)
    return P.ndim == 2 and not np.array_equal(P, np.diag(np.diag(P)))


def _diag_of(P, n) -> Optional[np.ndarray]:
    if P is None:
        return None
    P = np.asarray(P, dtype=np.float64)
    if P.ndim == 2:
        if _is_full_matrix(P):
            raise NotImplementedError("a full matrix P is handled by _matrix_p_factor_solver")
        P = np.diag(P)
    if P.shape != (n,):
        raise ValueError("solver[\"P\"] has the wrong size")
    # pinv of a diagonal matrix (`src/lsq.jl:457`)
    return np.where(P != 0, 1.0 / np.where(P != 0, P, 1.0), 0.0)


def lsqfit(db: LsqDB, *, solver: Optional[dict] = None, verbose=False, Ibasis=None, Itrain=None, E0=0.0,
           Vref=None, weights=None, sigmas=None, regularisers=(), deldb=False, error_table=False,
           saveqr: Optional[dict] = None, comm: "parallel.Communicator" = None):
    """`lsqfit(db; kwargs...) -> IP, fitinfo` (`src/lsq.jl:404-614`)."""
    solver = {"solver": "qr"} if solver is None else solver
    if weights is not None and sigmas is not None:
        raise ValueError("cannot specify both weights and sigmas")
    if sigmas is not None:
        # inverted in place in the caller's dict, as the reference does (`:424-429`)
        weights = sigmas
        for _cfg, cw in weights.items():
            if isinstance(cw, dict):
                for k in cw:
                    cw[k] = 1.0 / cw[k]
    weights = _fix_weights(weights)
    name = str(solver.get("solver", "qr")).lstrip(":")
    factor_solver = None
    if name == "svd":
        if "svd_ndiscard" not in solver:
            raise AssertionError('haskey(solver, "svd_ndiscard")')       # `@assert`, src/lsq.jl:483
        factor_solver = _svd_factor_solver(int(solver["svd_ndiscard"]))
    elif name == "rrqr":
        if "rrqr_tol" not in solver:
            raise AssertionError('haskey(solver, "rrqr_tol")')           # `@assert`, src/lsq.jl:490
        factor_solver = _rrqr_factor_solver(float(solver["rrqr_tol"]))
    elif name != "qr":
        raise ValueError("unknown `solver` in `lsqfit`" if name not in ("lsqr", "brr", "ard", "xgboost")
                         else f"solver :{name} is not provided by the CUDA path (:qr, :svd, :rrqr)")
    ctx = db._context()
    Y, W, Icfg = collect_observations(db, weights, Vref)
    Irows, Icols = _select(db, Y, W, Icfg, Ibasis, Itrain)
    n = len(db.basis) if Icols is None else len(Icols)
    if solver.get("P") is not None and _is_full_matrix(solver["P"]):
        if factor_solver is not None:
            raise NotImplementedError("a full matrix solver[\"P\"] is supported with the :qr solver (a diagonal P with all)")
        Pm = np.asarray(solver["P"], dtype=np.float64)
        if Pm.shape != (n, n):
            raise ValueError("solver[\"P\"] has the wrong size")
        Dinv = None
        factor_solver = _matrix_p_factor_solver(np.linalg.pinv(Pm))         # `D_inv = pinv(P)`, src/lsq.jl:457
    else:
        Dinv = _diag_of(solver.get("P"), n)
    if "P" in solver:
        del solver["P"]          # `src/lsq.jl:460-462`
    precon = _precon_blocks(db, weights, Irows)
    reg = _regulariser_rows(regularisers, db.basis, Icols) if regularisers else None
    # every row selected (`_select` returns sorted unique indices): no index upload, no gather on the device
    Irows_dev = None if len(Irows) == len(W) else Irows
    c, R, info = solve_device(ctx, db.Psi_dev, Y, W, Irows_dev, Icols, Dinv, want_R=True, comm=comm, precon=precon,
                              factor_solver=factor_solver, reg=reg, exchange=str(solver.get("exchange", "gram")))
    # `cond(qrPsi.R)` (`src/lsq.jl:469`; computed but only logged there): the device estimate (a lower bound, tight to a few
    # per cent) unless the exact SVD of the n x n factor is asked for (`verbose=True` or solver["exact_cond"])
    if "R_factor" in info:
        R = info.pop("R_factor")            # the factor of Psi pinv(P), which is what the reference keeps and conditions
        info.pop("cond_R_estimate", None)
    if verbose or solver.get("exact_cond") or "cond_R_estimate" not in info:
        kappa = float(np.linalg.cond(R)) if R is not None and R.size else float("nan")
    else:
        kappa = float(info["cond_R_estimate"])
    if isinstance(saveqr, dict):
        saveqr["R"] = R
        # the right-hand side of the factored system: weighted observations, then the regulariser targets
        # (`vcat(Y, Yreg)`, `src/lsq.jl:163-181`).  Rows transformed by a force preconditioner stay on the device and appear
        # here untransformed; Q (M x n) is not materialised at all (`saveqr["Q"]` is not written).
        saveqr["Y"] = (W * Y)[Irows] if reg is None or reg[0] is None else np.concatenate([(W * Y)[Irows], reg[1]])
    # combine(db.basis, c) uses the FULL basis (`src/lsq.jl:603`): scatter a subset fit
    cfull = np.zeros(len(db.basis))
    if Icols is None:
        cfull[:] = c
    else:
        cfull[Icols] = c
    IP = NBodyIP(db.basis, cfull)
    if Vref is not None:
        IP = SumIP(Vref, IP)
    infodict = asm_fitinfo(db, IP, c, Icols, weights, Vref, E0, regularisers, verbose, error_table,
                           cfull=cfull, comm=comm)
    if deldb:
        # the reference drops the matrix before its error table, which re-evaluates the IP (`src/lsq.jl:642-648`);
        # here the table is Psi*c, so the device matrix must outlive it
        db.delete()
    infodict["cond_R"] = kappa
    infodict["rel_rms"] = info["rel_rms"]
    infodict["solve_info"] = info
    infodict.update(solver)
    return IP, infodict


def asm_fitinfo(db, IP, c, Icols, weights, Vref, E0, regularisers, verbose, error_table, *, cfull, comm=None):
    """`asm_fitinfo` (`src/lsq.jl:617-678`)."""
    Jbasis = list(range(len(db.basis))) if Icols is None else [int(k) for k in Icols]
    infodict = {"E0": E0, "Ibasis": Jbasis, "c": np.array(c), "dbpath": db.dbpath, "weights": weights,
                "regularisers": list(regularisers), "errors": {}, "IPFitting_version": "na"}
    if error_table:
        from .lsqerrors import db_errors
        rm, rel = db_errors(db, cfull, Vref, comm=comm)
        infodict["errors"] = {"rmse": rm, "relrmse": rel}
    return infodict
