"""K4-K7 parity: get_lsq_system / lsqfit / error table through the C-ABI against the
numpy+LAPACK restatement of the reference (oracle/lsq_oracle.py) on the oracle's own Psi."""
import numpy as np
import pytest

import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import _capi, synth
from ipfitting_jl_b200.lsq import solve_device
from ipfitting_jl_b200.lsq_db import LsqDB
import ipf_oracle as O
import lsq_oracle as LO
from util import PSI_RTOL, psi_errors, rows_for, si_desc

pytestmark = pytest.mark.gpu

C_RTOL = 1e-8        # north_star: coefficients and per-configtype RMSEs agree to 1e-8


@pytest.fixture(scope="module")
def fitcase(ctx):
    sw = synth.StillingerWeber()
    cfgs = synth.rattled_configs("Si", 2, 12, seed=21, calc=sw, configtype="bulk")
    more = synth.rattled_configs("Si", 1, 10, seed=22, calc=sw, configtype="small")
    configs = cfgs + more
    # a weakly collinear basis so that coefficient parity at 1e-8 is meaningful (cond ~ 1e4)
    B = ipf.nbpolys(2, si_desc(2.0), 4) + ipf.nbpolys(3, si_desc(1.7), 2)
    db = LsqDB("", B, configs, ctx=ctx)
    r = rows_for(configs)
    ref = O.assemble(db.basis, configs, r["E"], r["F"], r["V"], db.nrows)
    return db, ref


def test_get_lsq_system_matches_hand_assembly(fitcase):
    """`test/test_lsq.jl:31-59`: the system equals w * vec_obs(evalb(B, at)) assembled by hand."""
    db, ref = fitcase
    w = {"default": {"E": 1.345, "F": 2.987}}
    A, Y = ipf.get_lsq_system(db, Vref=ipf.OneBody(0.0), weights=w)
    rowsel = []
    Aman, Yman = [], []
    for d in db.configs:
        nat = len(d.at)
        for okey, wt in (("E", 1.345 / np.sqrt(nat)), ("F", 2.987)):
            f, n = d.rows[okey]
            rowsel.append((f, okey))
            Aman.append(wt * ref[f:f + n, :]); Yman.append(wt * d.D[okey])
    order = np.argsort([f for f, _ in rowsel], kind="stable")
    Aman = np.vstack([Aman[k] for k in order]); Yman = np.concatenate([Yman[k] for k in order])
    assert A.shape == Aman.shape                 # V rows dropped: no "V" weight => w = 0 (Q1)
    assert np.allclose(Y, Yman, rtol=1e-13, atol=0)
    scale = np.abs(Aman).max(axis=0)
    assert (np.abs(A - Aman).max(axis=0) <= 1e-10 * scale).all()
    A2, Y2 = LO.get_lsq_system(ref, db.configs, weights=w, E0=0.0)
    assert np.allclose(Y, Y2, rtol=1e-13, atol=0) and A.shape == A2.shape


@pytest.mark.parametrize("weights", [None,
                                     {"default": {"E": 100.0, "F": 1.0, "V": 0.5}},
                                     {"default": {"E": 30.0, "F": 1.0, "V": 0.3}, "small": {"E": 10.0, "F": 2.0}},
                                     {"default": {"E": 10.0, "F": 1.0, "V": 1.0}, "ignore": ["small"]}])
def test_lsqfit_matches_reference_qr(fitcase, weights):
    db, ref = fitcase
    E0 = -4.0
    IP, info = ipf.lsqfit(db, weights=None if weights is None else dict(weights), E0=E0, Vref=ipf.OneBody(E0),
                          error_table=True)
    c_ref, kappa, rel_rms, R = LO.lsqfit_qr(ref, db.configs, weights, E0=E0)
    c = info["c"]
    assert kappa < 1e7
    assert np.linalg.norm(c - c_ref) <= C_RTOL * np.linalg.norm(c_ref)
    assert abs(info["rel_rms"] - rel_rms) <= 1e-8 * rel_rms
    assert abs(info["cond_R"] - kappa) <= 1e-6 * kappa
    err_ref, rel_ref = LO.rmse_table(ref, db.configs, c_ref, E0=E0)
    for ct in err_ref:
        for okey in err_ref[ct]:
            assert abs(info["errors"]["rmse"][ct][okey] - err_ref[ct][okey]) <= C_RTOL * err_ref[ct][okey]
            assert abs(info["errors"]["relrmse"][ct][okey] - rel_ref[ct][okey]) <= C_RTOL * rel_ref[ct][okey]


def test_subset_fit_and_column_scaling(fitcase):
    db, ref = fitcase
    Ib = ipf.filter_basis(db, lambda b: ipf.bodyorder(b) < 3)
    Itrain = list(range(0, len(db.configs), 2))
    P = np.linspace(0.5, 3.0, len(Ib))
    IP, info = ipf.lsqfit(db, Ibasis=Ib, Itrain=Itrain, solver={"solver": "qr", "P": P.copy()},
                          weights={"default": {"E": 50.0, "F": 1.0, "V": 1.0}})
    c_ref, kappa, rel_rms, _ = LO.lsqfit_qr(ref, db.configs, {"default": {"E": 50.0, "F": 1.0, "V": 1.0}},
                                            E0=None, Ibasis=Ib, Itrain=Itrain, P=P)
    assert np.linalg.norm(info["c"] - c_ref) <= C_RTOL * np.linalg.norm(c_ref)
    assert info["Ibasis"] == [int(k) for k in Ib]
    assert len(IP.c) == len(db.basis) and np.all(IP.c[len(Ib):] == 0)     # Q8: scattered into the full basis
    assert abs(info["rel_rms"] - rel_rms) <= 1e-8 * rel_rms


def test_ill_conditioned_system_needs_reorthogonalisation(ctx):
    """cond ~ 1e11: plain normal equations are useless; the shifted/iterated CholeskyQR
    must reproduce the Householder solution to the accuracy the conditioning allows."""
    rng = np.random.default_rng(5)
    M, n = 6000, 40
    U, _ = np.linalg.qr(rng.standard_normal((M, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    s = np.logspace(0, -11, n)
    A = np.asfortranarray((U * s) @ V.T)
    x = rng.standard_normal(n)
    y = A @ x + 1e-6 * rng.standard_normal(M)
    Pm = _capi.DeviceMatrix(ctx, M, n)
    Pm.from_host(A)
    c, R, info = solve_device(ctx, Pm, y, np.ones(M))
    import scipy.linalg as sla
    Q, Rr = sla.qr(A, mode="economic")
    c_ref = sla.solve_triangular(Rr, Q.T @ y)
    kappa = np.linalg.cond(Rr)
    # the first Gram matrix is numerically singular: at least one re-orthogonalisation pass, then either a third Gram
    # pass or the corrected semi-normal equations on the second
    assert info["passes"] >= 3 or (info["passes"] == 2 and info["csne_steps"] >= 1)
    # both are backward stable: compare residuals and the well-determined part of c
    r, r_ref = np.linalg.norm(A @ c - y), np.linalg.norm(A @ c_ref - y)
    assert abs(r - r_ref) <= 1e-8 * r_ref
    assert np.linalg.norm(A @ (c - c_ref)) <= 1e-9 * np.linalg.norm(y)
    assert abs(np.linalg.cond(R) - kappa) <= 1e-3 * kappa
    Pm.close()


def test_nan_checks(fitcase):
    db, _ = fitcase
    bad = db.configs[0].D["E"].copy()
    db.configs[0].D["E"][0] = np.nan
    try:
        with pytest.raises(_capi.IPFError, match="NaN detected in observations vector"):
            ipf.lsqfit(db)
    finally:
        db.configs[0].D["E"][:] = bad
    ctx = db._context()
    M = _capi.DeviceMatrix(ctx, 8, 2)
    A = np.ones((8, 2), order="F"); A[3, 1] = np.nan
    M.from_host(A)
    with pytest.raises(_capi.IPFError, match="discovered NaNs in LSQ system matrix"):
        solve_device(ctx, M, np.ones(8), np.ones(8))
    with pytest.raises(_capi.IPFError, match="NaN detected in weights vector"):
        w = np.ones(8); w[2] = np.nan
        M.from_host(np.ones((8, 2), order="F"))
        solve_device(ctx, M, np.ones(8), w)


def test_add_fits_equals_error_table(fitcase):
    """`test/test_errors.jl:54-71`: add_fits! + rmse(configs) == fitinfo["errors"]."""
    db, _ = fitcase
    E0 = -4.0
    IP, info = ipf.lsqfit(db, E0=E0, Vref=ipf.OneBody(E0), error_table=True,
                          weights={"default": {"E": 30.0, "F": 1.0, "V": 0.3}})
    ipf.add_fits(IP, db.configs, fitkey="IP", db=db)
    rm, rel = ipf.rmse(db.configs, fitkey="IP")
    for ct in rm:
        for okey in rm[ct]:
            assert abs(rm[ct][okey] - info["errors"]["rmse"][ct][okey]) <= 1e-9 * rm[ct][okey]
            assert abs(rel[ct][okey] - info["errors"]["relrmse"][ct][okey]) <= 1e-9 * rel[ct][okey]


def test_bench_basis_fit_matches_householder(ctx):
    """The bench basis (2B deg 14 + 3B deg 11, cond(R) ~ 1e12): the CholeskyQR solve must sit at the same
    least-squares minimum as LAPACK's Householder QR on the oracle's matrix.  The coefficient vector itself
    is only determined to cond*eps, so the comparison is on the residual and on the predictions."""
    import bench
    basis = bench.make_basis()
    configs = synth.rattled_configs("Si", 2, 40, seed=77, calc=synth.StillingerWeber())
    db = LsqDB("", basis, configs, ctx=ctx)
    r = rows_for(configs)
    ref = O.assemble(db.basis, configs, r["E"], r["F"], r["V"], db.nrows)
    w = {"default": {"E": 100.0, "F": 1.0, "V": 1.0}}
    IP, info = ipf.lsqfit(db, weights=dict(w), E0=-4.0, Vref=ipf.OneBody(-4.0))
    c_ref, kappa, rel_rms, _ = LO.lsqfit_qr(ref, configs, w, E0=-4.0)
    A, y = LO.get_lsq_system(ref, configs, w, E0=-4.0)
    assert kappa > 1e9
    # two backward-stable solvers agree on the minimum only to ~cond*eps (1e-4 here); observed ~5e-8
    assert abs(info["rel_rms"] - rel_rms) <= 1e-6 * rel_rms
    r_gpu = np.linalg.norm(A @ info["c"] - y) / np.linalg.norm(y)
    assert abs(r_gpu - rel_rms) <= 1e-6 * rel_rms                     # the returned c reproduces the minimum
    assert np.linalg.norm(A @ (info["c"] - c_ref)) <= 2e-3 * np.linalg.norm(y) * rel_rms
    assert abs(np.log10(info["cond_R"]) - np.log10(kappa)) < 0.5
    # per-configtype RMSE table of this fit against the oracle's table for ITS minimiser: both sit at the same minimum of a
    # cond ~ 1e12 system, where the coefficient vectors differ by ~cond*eps in the flat directions: the tables agree to
    # ~1e-6 relative (observed 4e-8 .. 1.3e-6 depending on the rounding of Psi), not to the 1e-8 of the well-conditioned cases
    IP2, info2 = ipf.lsqfit(db, weights=dict(w), E0=-4.0, Vref=ipf.OneBody(-4.0), error_table=True)
    rm_ref, _ = LO.rmse_table(ref, configs, c_ref, E0=-4.0)
    for ct in rm_ref:
        for okey in rm_ref[ct]:
            assert abs(info2["errors"]["rmse"][ct][okey] - rm_ref[ct][okey]) <= 1e-5 * rm_ref[ct][okey], (ct, okey)


@pytest.mark.parametrize("ndiscard", [0, 2])
def test_lsqfit_svd_matches_reference(fitcase, ndiscard):
    """`:svd` (`src/lsq.jl:482-488`): SVD of the n x n factor on the host, Q'y and the residual from the device."""
    db, ref = fitcase
    w = {"default": {"E": 30.0, "F": 1.0, "V": 0.5}}
    E0 = -4.0
    IP, info = ipf.lsqfit(db, weights=dict(w), E0=E0, Vref=ipf.OneBody(E0),
                          solver={"solver": "svd", "svd_ndiscard": ndiscard})
    c_ref, rel_rms, S = LO.lsqfit_svd(ref, db.configs, ndiscard, w, E0=E0)
    assert np.linalg.norm(info["c"] - c_ref) <= C_RTOL * np.linalg.norm(c_ref)
    assert abs(info["rel_rms"] - rel_rms) <= 1e-8 * rel_rms
    assert np.allclose(info["solve_info"]["svd_S"], S, rtol=1e-9, atol=0)
    if ndiscard == 0:            # full rank: the same minimiser as :qr
        c_qr = LO.lsqfit_qr(ref, db.configs, w, E0=E0)[0]
        assert np.linalg.norm(info["c"] - c_qr) <= 1e-7 * np.linalg.norm(c_qr)
    with pytest.raises(AssertionError):
        ipf.lsqfit(db, weights=dict(w), solver={"solver": "svd"})


@pytest.mark.parametrize("tol", [1e-14, 3e-3])
def test_lsqfit_rrqr_matches_reference(fitcase, tol):
    """`:rrqr` (`src/lsq.jl:489-496`): column-pivoted QR of the n x n factor selects the same columns as that of Psi."""
    db, ref = fitcase
    w = {"default": {"E": 30.0, "F": 1.0, "V": 0.5}}
    P = np.linspace(0.7, 2.0, len(db.basis))
    IP, info = ipf.lsqfit(db, weights=dict(w), E0=0.0, solver={"solver": "rrqr", "rrqr_tol": tol, "P": P.copy()})
    c_ref, rel_rms, rank = LO.lsqfit_rrqr(ref, db.configs, tol, w, E0=0.0, P=P)
    assert info["solve_info"]["rrqr_rank"] == rank
    assert (rank < len(db.basis)) == (tol > 1e-6)
    assert np.array_equal(info["c"] == 0, c_ref == 0)            # the same columns were discarded
    assert np.linalg.norm(info["c"] - c_ref) <= C_RTOL * np.linalg.norm(c_ref)
    assert abs(info["rel_rms"] - rel_rms) <= 1e-8 * rel_rms


def test_regularisers_append_rows(fitcase):
    """`_regularise!` (`src/lsq.jl:163-181`): matrices (target 0) and (P, q) pairs are stacked below the weighted
    system; the solve sees them as ordinary rows (also after the column scaling `solver["P"]`)."""
    db, ref = fitcase
    n = len(db.basis)
    rng = np.random.default_rng(5)
    w = {"default": {"E": 30.0, "F": 1.0, "V": 0.5}}
    regs = [0.3 * np.eye(n), (rng.normal(size=(3, n)), rng.normal(size=3))]
    A, Y = ipf.get_lsq_system(db, Vref=ipf.OneBody(0.0), weights=dict(w), regularisers=regs)
    A2, Y2 = LO.get_lsq_system(ref, db.configs, weights=w, E0=0.0, regularisers=regs)
    assert A.shape == A2.shape and np.array_equal(A[-(n + 3):], A2[-(n + 3):]) and np.array_equal(Y[-(n + 3):], Y2[-(n + 3):])
    P = np.linspace(0.7, 2.0, n)
    IP, info = ipf.lsqfit(db, weights=dict(w), E0=0.0, regularisers=regs, solver={"solver": "qr", "P": P.copy()})
    c_ref, kappa, rel_rms, _ = LO.lsqfit_qr(ref, db.configs, w, E0=0.0, P=P, regularisers=regs)
    assert np.linalg.norm(info["c"] - c_ref) <= C_RTOL * np.linalg.norm(c_ref)
    assert abs(info["rel_rms"] - rel_rms) <= 1e-8 * rel_rms
    c_plain = LO.lsqfit_qr(ref, db.configs, w, E0=0.0, P=P)[0]
    assert np.linalg.norm(c_ref - c_plain) > 1e-3 * np.linalg.norm(c_plain)      # the regularisers do act


def test_vref_from_a_previous_fit_is_evaluated_on_the_device(fitcase, ctx):
    """`Vref` = a fitted potential (`eval_obs(okey, Vref, dat)`, `src/lsq.jl:96-99`): the 2-body fit is subtracted from
    every observation before the 3-body fit; evaluated as Psi_ref c_ref for all configurations at once."""
    db, ref = fitcase
    w = {"default": {"E": 30.0, "F": 1.0, "V": 0.5}}
    B2 = ipf.nbpolys(2, si_desc(2.0), 4)
    db2 = LsqDB("", B2, db.configs, ctx=ctx)
    IP2, info2 = ipf.lsqfit(db2, weights=dict(w), E0=-4.0, Vref=ipf.OneBody(-4.0))       # SumIP(OneBody, NBodyIP)
    from ipfitting_jl_b200.lsq import collect_observations, _fix_weights
    Y, W, Icfg = collect_observations(db, _fix_weights(dict(w)), IP2)
    # by hand: data - (Psi_2 c_2 + E0 N on the energy rows)
    yref = db2.Psi @ info2["c"]
    Ydata = np.zeros(db.nrows)
    for d in db.configs:
        for okey, v in d.D.items():
            f, n = d.rows[okey]
            Ydata[f:f + n] = v
            if okey == "E":
                yref[f] += -4.0 * len(d.at)
    assert np.allclose(Y, Ydata - yref, rtol=1e-12, atol=1e-9 * np.abs(Ydata).max())
    # and the observation-by-observation path of the reference gives the same numbers
    db.configs[0].info["W"] = {"E": 30.0, "F": 1.0, "V": 0.5}      # forces the loop
    try:
        Yl, Wl, _ = collect_observations(db, _fix_weights(dict(w)), IP2)
    finally:
        del db.configs[0].info["W"]
    assert np.allclose(Yl, Y, rtol=1e-10, atol=1e-9 * np.abs(Ydata).max())
    IP3, info3 = ipf.lsqfit(db, weights=dict(w), Vref=IP2)
    assert info3["rel_rms"] < 1.0 and np.isfinite(info3["c"]).all()


def test_species_fit_on_mixed_size_cells(ctx):
    """BASELINE config 5 in miniature: two-species (Si/C zincblende) cells of 8 to 512 atoms, species-resolved 2- and 3-body
    basis, labels generated from a known coefficient vector through the ORACLE's design matrix: `LsqDB` + `lsqfit` on the
    GPU must recover the coefficients (the rows of the 512-atom cells go through the large-configuration paths)."""
    desc = ipf.BondAngleDesc("exp(- (r/1.89 - 1.0))", ipf.CosCut(3.2, 4.2))
    B = ipf.NBodyBasis(ipf.species_nbpolys(2, desc, 4, [14, 6]) + ipf.species_nbpolys(3, desc, 2, [14, 6]))
    rng = np.random.default_rng(23)
    ats = []
    for L in (1, 1, 2, 1, 3, 2, 4, 1, 2, 1):
        ats.append(synth.rattle(synth.repeat(synth.zincblende(4.36, 14, 6), L), 0.1 + 0.1 * rng.uniform(), rng, 0.03))
    blank = [ipf.Dat(a, "sic", E=0.0, F=np.zeros((len(a), 3)), V=np.zeros(6)) for a in ats]
    from ipfitting_jl_b200.lsq_db import _alloc_rows
    nrows = _alloc_rows(blank)
    r = rows_for(blank)
    ref = O.assemble(B, blank, r["E"], r["F"], r["V"], nrows)
    c_true = rng.standard_normal(len(B)) / (1.0 + np.arange(len(B)))
    y = ref @ c_true
    configs = []
    for d in blank:
        e = y[d.rows["E"][0]]
        f = y[d.rows["F"][0]:d.rows["F"][0] + d.rows["F"][1]].reshape(-1, 3)
        v = y[d.rows["V"][0]:d.rows["V"][0] + 6]
        configs.append(ipf.Dat(d.at, "sic", E=e, F=f, V=v))
    db = LsqDB("", B, configs, ctx=ctx)
    assert psi_errors(db.Psi, ref, configs) <= PSI_RTOL
    w = {"default": {"E": 10.0, "F": 1.0, "V": 1.0}}
    IP, info = ipf.lsqfit(db, weights=dict(w), E0=0.0, error_table=True)
    kappa = info["cond_R"]
    assert np.linalg.norm(info["c"] - c_true) <= max(1e-8, 100 * kappa * 2.2e-16) * np.linalg.norm(c_true)
    assert info["rel_rms"] <= 1e-9
    assert max(len(a) for a in ats) == 512


def test_cond_estimate_is_a_tight_lower_bound(fitcase):
    """`cond(qrPsi.R)` (`src/lsq.jl:469`): the device estimate (power / inverse iteration on the returned factor) is a lower
    bound of the SVD value and within 20 % of it."""
    db, _ = fitcase
    w = {"default": {"E": 30.0, "F": 1.0, "V": 0.5}}
    _, est = ipf.lsqfit(db, weights=dict(w), E0=-4.0)
    _, exact = ipf.lsqfit(db, weights=dict(w), E0=-4.0, solver={"solver": "qr", "exact_cond": True})
    assert est["cond_R"] <= exact["cond_R"] * (1.0 + 1e-9)
    assert est["cond_R"] >= 0.8 * exact["cond_R"]


@pytest.mark.parametrize("n", [150, 216, 256, 257, 400])
def test_solve_sizes_around_the_kernel_switches(ctx, n):
    """Column counts on both sides of the single-CTA / multi-CTA Cholesky switch (n = 256) and of the shared-memory opt-in
    of the single-CTA kernel (n ~ 170): coefficients against LAPACK on a moderately conditioned random system."""
    rng = np.random.default_rng(n)
    M = 4 * n + 37
    A = np.asfortranarray(rng.standard_normal((M, n)) * np.logspace(0, -3, n)[None, :])
    x = rng.standard_normal(n)
    y = A @ x + 1e-3 * rng.standard_normal(M)
    Pm = _capi.DeviceMatrix(ctx, M, n)
    Pm.from_host(A)
    c, R, info = solve_device(ctx, Pm, y, np.ones(M))
    c_ref = np.linalg.lstsq(A, y, rcond=None)[0]
    assert np.linalg.norm(c - c_ref) <= 1e-9 * np.linalg.norm(c_ref)
    Pm.close()


def test_env_fit_with_configtype_weights(ctx):
    """BASELINE config 4 in miniature: W bcc cells with and without vacancies (configtypes "rand" / "vacancy"), an
    environment-dependent 3- and 4-body basis, per-configtype weights (`src/lsq.jl:137-146`); labels generated from a known
    coefficient vector through the ORACLE's matrix: the weighted fit recovers it and the error table has one entry per
    configtype."""
    r0 = ipf.rnn("W")
    d3 = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(1.5 * r0 - 0.6, 1.5 * r0))
    d4 = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(1.25 * r0 - 0.4, 1.25 * r0))
    B = ipf.NBodyBasis(ipf.nbpolys(2, d3, 6) + ipf.envpolys(3, d3, 2, 1, ipf.CosCut(1.1 * r0, 1.45 * r0))
                       + ipf.envpolys(4, d4, 2, 1, ipf.CosCut(1.05 * r0, 1.2 * r0)))
    blank = synth.rattled_configs("W", 2, 12, seed=41, calc=synth.PairPotential(r0), vacancies=2)
    assert {d.configtype for d in blank} == {"rand", "vacancy"}
    from ipfitting_jl_b200.lsq_db import _alloc_rows
    nrows = _alloc_rows(blank)
    r = rows_for(blank)
    ref = O.assemble(B, blank, r["E"], r["F"], r["V"], nrows)
    rng = np.random.default_rng(5)
    c_true = rng.standard_normal(len(B)) / (1.0 + np.arange(len(B)))
    y = ref @ c_true
    configs = []
    for d in blank:
        configs.append(ipf.Dat(d.at, d.configtype, E=y[d.rows["E"][0]],
                               F=y[d.rows["F"][0]:d.rows["F"][0] + d.rows["F"][1]].reshape(-1, 3),
                               V=y[d.rows["V"][0]:d.rows["V"][0] + 6]))
    db = LsqDB("", B, configs, ctx=ctx)
    assert psi_errors(db.Psi, ref, configs) <= PSI_RTOL
    w = {"default": {"E": 10.0, "F": 1.0, "V": 1.0}, "vacancy": {"E": 40.0, "F": 3.0, "V": 0.5}}
    IP, info = ipf.lsqfit(db, weights={k: dict(v) for k, v in w.items()}, E0=0.0, error_table=True)
    assert np.linalg.norm(info["c"] - c_true) <= max(1e-8, 100 * info["cond_R"] * 2.2e-16) * np.linalg.norm(c_true)
    assert set(info["errors"]["rmse"]) == {"rand", "vacancy", "set"}
    c_ref = LO.lsqfit_qr(ref, configs, w, E0=0.0)[0]
    assert np.linalg.norm(info["c"] - c_ref) <= max(1e-8, 100 * info["cond_R"] * 2.2e-16) * np.linalg.norm(c_ref)
