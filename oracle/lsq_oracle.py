"""numpy/LAPACK restatement of the reference's lsqfit path (TEST INFRASTRUCTURE ONLY;
parity unpinned -- see oracle/ipf_oracle.c).  Follows, line by line:

  collect_observations   src/lsq.jl:74-109      _get_weights  src/lsq.jl:118-161
  get_lsq_system         src/lsq.jl:245-311     lsqfit :qr    src/lsq.jl:450-473,593
  _fiterrs (rmse)        src/lsqerrors.jl:225-266
  weighthook/err_weighthook  src/datatypes.jl:29-30,66-67; src/prototypes.jl:125,133

Householder QR is LAPACK dgeqrf through scipy -- the routine Julia's `qr!` calls.
Works on plain dict/array inputs so that it shares no code with the product's host layer.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla


def _wh(okey, nat):
    return 1.0 / np.sqrt(nat) if okey in ("E", "V") else 1.0


def _ewh(okey, nat):
    return 1.0 / nat if okey in ("E", "V") else 1.0


def observations(configs):
    for icfg, d in enumerate(configs):
        for okey in list(d.D.keys()):
            yield okey, d, icfg


def rows_of(d, okey):
    f, n = d.rows[okey]
    return np.arange(f, f + n)


def vref_obs(E0, okey, d):
    """OneBody(E0): E0*nat on E rows, 0 on F and V rows."""
    n = len(d.D[okey])
    if okey == "E":
        return np.array([E0 * len(d.at)])
    return np.zeros(n)


def collect_observations(configs, nrows, weights, E0):
    Y = np.zeros(nrows); W = np.zeros(nrows); Icfg = np.zeros(nrows, dtype=np.int64) - 1
    for okey, d, icfg in observations(configs):
        if d.configtype in weights["ignore"] or okey in weights["ignore"]:
            continue
        rows = rows_of(d, okey)
        obs = d.D[okey].copy()
        if "Vref" in d.info:
            obs = obs - np.asarray(d.info["Vref"][okey])
        elif E0 is not None:
            obs = obs - vref_obs(E0, okey, d)
        Y[rows] = obs
        Icfg[rows] = icfg
        # _get_weights
        w = 0.0
        if "W" in d.info:
            if okey in d.info["W"]:
                w = d.info["W"][okey]
        else:
            key = d.configtype if d.configtype in weights else "default"
            if okey in weights[key]:
                w = weights[key][okey] * _wh(okey, len(d.at))
        W[rows] = np.broadcast_to(np.asarray(w, dtype=np.float64), (len(rows),))
    return Y, W, Icfg


def fix_weights(weights):
    weights = dict(weights or {})
    weights.setdefault("ignore", [])
    weights.setdefault("default", {"E": 1.0, "F": 1.0, "V": 1.0})
    return weights


def get_lsq_system(Psi, configs, weights=None, E0=0.0, Ibasis=None, Itrain=None, regularisers=()):
    weights = fix_weights(weights)
    Y, W, Icfg = collect_observations(configs, Psi.shape[0], weights, E0)
    assert not np.isnan(Y).any() and not np.isnan(W).any()
    keep = W != 0
    if Itrain is not None:
        keep &= np.isin(Icfg, np.asarray(list(Itrain)))
    Irows = np.nonzero(keep)[0]
    A = Psi[Irows, :] if Ibasis is None else Psi[np.ix_(Irows, np.asarray(list(Ibasis)))]
    A = np.array(A, order="F", copy=True)
    Y = Y[Irows]; W = W[Irows]
    assert not np.isnan(A).any()
    A, Y = W[:, None] * A, W * Y
    # `_regularise!` (`src/lsq.jl:163-181`): vcat below the weighted system; a bare matrix has target zero
    for reg in regularisers:
        P, q = reg if isinstance(reg, tuple) else (reg, None)
        P = np.asarray(P, dtype=float)
        if Ibasis is not None:
            P = P[:, np.asarray(list(Ibasis))]
        A = np.vstack([A, P]); Y = np.concatenate([Y, np.zeros(P.shape[0]) if q is None else np.asarray(q, dtype=float)])
    return A, Y


def lsqfit_qr(Psi, configs, weights=None, E0=None, Ibasis=None, Itrain=None, P=None, regularisers=()):
    """Returns c, cond(R), rel_rms, R exactly as the :qr branch computes them."""
    A, Y = get_lsq_system(Psi, configs, weights, E0, Ibasis, Itrain, regularisers)
    n = A.shape[1]
    if P is not None and np.ndim(P) == 2:
        # a full matrix (the "Laplacian preconditioning" of `src/lsq.jl:450-458`): D_inv = pinv(P); mul!(Psi, Psi, D_inv)
        Dm = np.linalg.pinv(np.asarray(P, dtype=float))
        A = A @ Dm
        Q, R = sla.qr(A, mode="economic")
        kappa = np.linalg.cond(R)
        c = sla.solve_triangular(R, Q.T @ Y)
        rel_rms = np.linalg.norm(Q @ (R @ c) - Y) / np.linalg.norm(Y)
        return Dm @ c, kappa, rel_rms, R                     # `c = D_inv * c`, src/lsq.jl:593
    Dinv = np.ones(n) if P is None else np.where(np.asarray(P) != 0, 1.0 / np.asarray(P, dtype=float), 0.0)
    A = A * Dinv[None, :]
    Q, R = sla.qr(A, mode="economic")
    kappa = np.linalg.cond(R)
    c = sla.solve_triangular(R, Q.T @ Y)
    rel_rms = np.linalg.norm(Q @ (R @ c) - Y) / np.linalg.norm(Y)
    return Dinv * c, kappa, rel_rms, R


def lsqfit_svd(Psi, configs, svd_ndiscard, weights=None, E0=None, Ibasis=None, Itrain=None, P=None):
    """The :svd branch (`src/lsq.jl:482-488`): F = svd(Psi);
    c = F.V[:, 1:end-k] * (Diagonal(F.S[1:end-k]) \\ (F.U' * Y)[1:end-k]); rel_rms = norm(Psi*c - Y) / norm(Y)."""
    A, Y = get_lsq_system(Psi, configs, weights, E0, Ibasis, Itrain)
    n = A.shape[1]
    Dinv = np.ones(n) if P is None else np.where(np.asarray(P) != 0, 1.0 / np.asarray(P, dtype=float), 0.0)
    A = A * Dinv[None, :]
    U, S, Vt = np.linalg.svd(A, full_matrices=False)
    k = n - svd_ndiscard
    c = Vt[:k].T @ ((U.T @ Y)[:k] / S[:k])
    rel_rms = np.linalg.norm(A @ c - Y) / np.linalg.norm(Y)
    return Dinv * c, rel_rms, S


def lsqfit_rrqr(Psi, configs, rrqr_tol, weights=None, E0=None, Ibasis=None, Itrain=None, P=None):
    """The :rrqr branch (`src/lsq.jl:489-496`): qrPsi = pqrfact(Psi, rtol=rrqr_tol); c = qrPsi \\ Y.  pqrfact
    (LowRankApprox.jl, not vendored) = column-pivoted Householder QR stopped where |R_kk| <= rtol |R_11|; `\\`
    gives the basic solution on the selected columns.  Returns c, rel_rms, rank."""
    A, Y = get_lsq_system(Psi, configs, weights, E0, Ibasis, Itrain)
    n = A.shape[1]
    Dinv = np.ones(n) if P is None else np.where(np.asarray(P) != 0, 1.0 / np.asarray(P, dtype=float), 0.0)
    A = A * Dinv[None, :]
    Q, R, piv = sla.qr(A, mode="economic", pivoting=True)
    d = np.abs(np.diag(R))
    k = int(np.sum(d > rrqr_tol * d[0]))
    c = np.zeros(n)
    c[piv[:k]] = sla.solve_triangular(R[:k, :k], (Q.T @ Y)[:k])
    rel_rms = np.linalg.norm(A @ c - Y) / np.linalg.norm(Y)
    return Dinv * c, rel_rms, k


def rmse_table(Psi, configs, cfull, E0=None):
    """_fiterrs with fit values Psi*c + Vref observations; raw observations, err_weighthook."""
    yhat = Psi @ cfull
    cts = []
    for d in configs:
        if d.configtype not in cts:
            cts.append(d.configtype)
    names = cts + ["set"]
    err = {k: {} for k in names}; nrm = {k: {} for k in names}; ln = {k: {} for k in names}
    for okey, d, _ in observations(configs):
        w = _ewh(okey, len(d.at))
        ex = d.D[okey] * w
        fit = yhat[rows_of(d, okey)] + (vref_obs(E0, okey, d) if E0 is not None else 0.0)
        fit = fit * w
        for ct in (d.configtype, "set"):
            err[ct][okey] = err[ct].get(okey, 0.0) + np.linalg.norm(ex - fit) ** 2
            nrm[ct][okey] = nrm[ct].get(okey, 0.0) + np.linalg.norm(ex) ** 2
            ln[ct][okey] = ln[ct].get(okey, 0.0) + len(ex)
    rel = {k: {} for k in names}
    for ct in names:
        for okey in err[ct]:
            err[ct][okey] = np.sqrt(err[ct][okey] / ln[ct][okey])
            nrm[ct][okey] = np.sqrt(nrm[ct][okey] / ln[ct][okey])
            rel[ct][okey] = err[ct][okey] / nrm[ct][okey]
    return err, rel


# ---------------------------------------------------------------- force preconditioner
def adjacency_precon(at, intervals, stab=1e-4, innerstab=0.1):
    """AdjacancyPrecon(at) restated loop by loop (src/preconditioners.jl:42-74); neighbour list from the
    C oracle's brute-force enumeration."""
    import ipf_oracle as O
    n = len(at)
    rc = max(i[1] for i in intervals)
    I, J, S, R = O.neighbours(at.X, at.cell, at.pbc, rc)
    P = np.zeros((3 * n, 3 * n))
    for i, j, Rij in zip(I, J, R):
        if j < i:
            continue
        r = np.linalg.norm(Rij)
        Rh = Rij / r
        Pi = (1 - innerstab) * np.outer(Rh, Rh) + innerstab * np.eye(3)
        Pij = 0.0 * Pi
        for (r0, r1, k) in intervals:
            if r0 <= r <= r1:
                Pij = k * Pi
                break
        for a in range(3):
            for b in range(3):
                P[3 * i + a, 3 * j + b] -= Pij[a, b]
                P[3 * j + b, 3 * i + a] -= Pij[a, b]
                P[3 * i + a, 3 * i + b] += Pij[a, b]
                P[3 * j + a, 3 * j + b] += Pij[a, b]
    return P + stab * np.eye(3 * n)


def forceprecon_system(Psi, configs, weights, precon_fn, E0=0.0, solver="chol"):
    """get_lsq_system with _forceprecon (src/lsq.jl:184-234, 245-311): rtP \\ on the F rows of Psi and Y
    BEFORE the weights are applied."""
    weights = fix_weights(weights)
    Y, W, Icfg = collect_observations(configs, Psi.shape[0], weights, E0)
    Psi = Psi.copy()
    for okey, d, _ in observations(configs):
        if okey != "F":
            continue
        P = precon_fn(d.at)
        if solver == "chol":
            rt = np.linalg.cholesky(P)
        else:
            rt = np.real(sla.sqrtm(P))
        rows = rows_of(d, okey)
        Y[rows] = np.linalg.solve(rt, Y[rows])
        Psi[rows, :] = np.linalg.solve(rt, Psi[rows, :])
    keep = np.nonzero(W != 0)[0]
    return W[keep, None] * Psi[keep, :], W[keep] * Y[keep]
