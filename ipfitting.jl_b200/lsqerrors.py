"""Fit errors: `add_fits!`, `rmse`, and the error table of `lsqfit(error_table=true)`.

Mirrors `/root/reference/src/lsqerrors.jl`:
  add_fits! / add_fits_serial!   :62-90   (store vec_obs(eval_obs(okey, IP, cfg)) in cfg.info[fitkey][okey])
  _fiterrs -> rmse / mae         :225-266 (per configtype and "set": sqrt(sum (w (y - yfit))^2 / len), w = err_weighthook)
For a potential that is linear in the basis of the db, the fit values are Psi*c
(+ the Vref observations), computed on the GPU (`ipf_predict`, `ipf_residual_stats`)
instead of re-evaluating the potential on every configuration (`src/lsq.jl:642-648`).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import numpy as np

from . import _capi
from .datatypes import err_weighthook
from .lsq_db import LsqDB, matrows
from .obsiter import observations
from .potentials import linear_part, vref_observation


def _vref_rows(db: LsqDB, others) -> np.ndarray:
    off = np.zeros(db.nrows)
    if others:
        for okey, dat, _ in observations(db):
            rows = matrows(dat, okey)
            for V in others:
                off[rows] += vref_observation(V, okey, dat)
    return off


def db_errors(db: LsqDB, cfull: np.ndarray, Vref=None, comm=None):
    """rmse, relrmse dicts {configtype | "set": {okey: value}} for the fit `cfull` on the
    configs of `db` (`Err.rmse(db.configs; fitkey="IP")`, `src/lsq.jl:645`)."""
    ctx = db._context()
    cts: List[str] = []
    for d in db.configs:
        if d.configtype not in cts:
            cts.append(d.configtype)
    okeys = ["E", "F", "V"]
    gid = {(ct, k): i * 3 + j for i, ct in enumerate(cts) for j, k in enumerate(okeys)}
    group = np.full(db.nrows, -1, dtype=np.int32)
    werr = np.zeros(db.nrows)
    Yraw = np.zeros(db.nrows)
    for okey, dat, _ in observations(db):
        rows = matrows(dat, okey)
        group[rows] = gid[(dat.configtype, okey)]
        werr[rows] = err_weighthook(okey, dat)
        Yraw[rows] = dat.D[okey]
    Yoff = _vref_rows(db, [Vref] if Vref is not None else [])
    ng = max(1, 3 * len(cts))
    sse, ssy, cnt = np.zeros(ng), np.zeros(ng), np.zeros(ng, dtype=np.int64)
    c = np.ascontiguousarray(cfull, dtype=np.float64)
    ctx.check(ctx.lib.ipf_residual_stats(
        ctx.h, db.Psi_dev.h, _capi.ptr(c, C.c_double), _capi.ptr(Yraw, C.c_double), _capi.ptr(Yoff, C.c_double),
        _capi.ptr(werr, C.c_double), _capi.ptr(group, C.c_int32), ng, _capi.ptr(sse, C.c_double),
        _capi.ptr(ssy, C.c_double), _capi.ptr(cnt, C.c_int64)))
    if comm is not None and comm.world_size > 1:
        # config types may differ between shards: reduce by name
        sse, ssy, cnt, cts = comm.allreduce_named_stats(cts, okeys, sse, ssy, cnt)
        gid = {(ct, k): i * 3 + j for i, ct in enumerate(cts) for j, k in enumerate(okeys)}
    return _finish_errors(cts, okeys, gid, sse, ssy, cnt)


def _finish_errors(cts, okeys, gid, sse, ssy, cnt):
    rm: Dict[str, Dict[str, float]] = {ct: {} for ct in cts}
    rel: Dict[str, Dict[str, float]] = {ct: {} for ct in cts}
    rm["set"], rel["set"] = {}, {}
    tot = {k: [0.0, 0.0, 0] for k in okeys}
    for ct in cts:
        for k in okeys:
            g = gid[(ct, k)]
            if cnt[g] == 0:
                continue
            e = np.sqrt(sse[g] / cnt[g])
            nrm = np.sqrt(ssy[g] / cnt[g])
            rm[ct][k] = float(e)
            rel[ct][k] = float(e / nrm) if nrm > 0 else float("inf")
            tot[k][0] += sse[g]; tot[k][1] += ssy[g]; tot[k][2] += int(cnt[g])
    for k in okeys:
        if tot[k][2] == 0:
            continue
        e = np.sqrt(tot[k][0] / tot[k][2])
        nrm = np.sqrt(tot[k][1] / tot[k][2])
        rm["set"][k] = float(e)
        rel["set"][k] = float(e / nrm) if nrm > 0 else float("inf")
    return rm, rel


def add_fits(IP, configs, fitkey: str = "fit", ctx=None, db: LsqDB = None):
    """`add_fits!(IP, configs; fitkey)` (`src/lsqerrors.jl:62-77`): stores the fitted observation
    values in `cfg.info[fitkey][okey]`.  A linear-in-basis IP is evaluated as Psi*c on the GPU
    (pass the `db` it was fitted on to reuse its matrix)."""
    lin, others = linear_part(IP)
    if lin is None:
        for okey, dat, _ in observations(configs):
            dat.info.setdefault(fitkey, {})[okey] = vref_observation(IP, okey, dat)
        return
    if db is None or db.configs is not configs and list(db.configs) != list(configs):
        db = LsqDB("", lin.basis, configs, ctx=ctx)
    c = np.ascontiguousarray(lin.c, dtype=np.float64)
    yhat = np.empty(db.nrows)
    cx = db._context()
    cx.check(cx.lib.ipf_predict(cx.h, db.Psi_dev.h, _capi.ptr(c, C.c_double), _capi.ptr(yhat, C.c_double)))
    yhat += _vref_rows(db, others)
    for okey, dat, _ in observations(db):
        dat.info.setdefault(fitkey, {})[okey] = yhat[matrows(dat, okey)].copy()


def _fiterrs(configs, fitkey="fit", p=2) -> Tuple[dict, dict]:
    """`_fiterrs` (`src/lsqerrors.jl:225-266`) on stored fit values (host-side analysis)."""
    cts = []
    for d in configs:
        if d.configtype not in cts:
            cts.append(d.configtype)
    names = cts + ["set"]
    err = {ct: {} for ct in names}; nrm = {ct: {} for ct in names}; ln = {ct: {} for ct in names}
    for okey, dat, _ in observations(configs):
        w = err_weighthook(okey, dat)
        ex = dat.D[okey] * w
        fit = np.asarray(dat.info[fitkey][okey]) * w
        for ct in (dat.configtype, "set"):
            err[ct][okey] = err[ct].get(okey, 0.0) + float(np.sum(np.abs(ex - fit) ** p))
            nrm[ct][okey] = nrm[ct].get(okey, 0.0) + float(np.sum(np.abs(ex) ** p))
            ln[ct][okey] = ln[ct].get(okey, 0.0) + len(ex)
    rel = {ct: {} for ct in names}
    for ct in names:
        for okey in err[ct]:
            err[ct][okey] = (err[ct][okey] / ln[ct][okey]) ** (1.0 / p)
            nrm[ct][okey] = (nrm[ct][okey] / ln[ct][okey]) ** (1.0 / p)
            rel[ct][okey] = err[ct][okey] / nrm[ct][okey]
    return err, rel


def rmse(configs, fitkey="fit"):
    return _fiterrs(configs, fitkey, 2)


def mae(configs, fitkey="fit"):
    return _fiterrs(configs, fitkey, 1)
