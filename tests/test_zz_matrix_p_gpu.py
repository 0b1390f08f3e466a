"""`solver["P"]` as a full matrix through the C-ABI against the numpy+LAPACK restatement of the reference
(oracle/lsq_oracle.py) on the oracle's own Psi.  (A file of its own, last in collection order.)"""
import numpy as np
import pytest

import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import synth
from ipfitting_jl_b200.lsq_db import LsqDB
import ipf_oracle as O
import lsq_oracle as LO
from util import rows_for, si_desc

pytestmark = pytest.mark.gpu

C_RTOL = 1e-8        # north_star: coefficients agree to 1e-8


@pytest.fixture(scope="module")
def fitcase(ctx):
    sw = synth.StillingerWeber()
    configs = synth.rattled_configs("Si", 2, 6, seed=21, calc=sw, configtype="bulk") + \
        synth.rattled_configs("Si", 1, 6, seed=22, calc=sw, configtype="small")
    B = ipf.nbpolys(2, si_desc(2.0), 4) + ipf.nbpolys(3, si_desc(1.7), 2)
    db = LsqDB("", B, configs, ctx=ctx)
    r = rows_for(configs)
    ref = O.assemble(db.basis, configs, r["E"], r["F"], r["V"], db.nrows)
    return db, ref


def test_full_matrix_column_preconditioner(fitcase):
    """`solver["P"]` as a full matrix (the reference's "Laplacian preconditioning", `src/lsq.jl:450-458,593`):
    Psi <- Psi pinv(P), QR, c = pinv(P) (qr \\ Y); R and cond(R) are those of the transformed matrix."""
    db, ref = fitcase
    n = len(db.basis)
    P = 2.5 * np.eye(n) - np.diag(np.ones(n - 1), 1) - np.diag(np.ones(n - 1), -1)
    P[0, n - 1] = 0.3                                   # not symmetric either
    w = {"default": {"E": 40.0, "F": 1.0, "V": 0.5}}
    saveqr = {}
    IP, info = ipf.lsqfit(db, weights=dict(w), E0=0.0, solver={"solver": "qr", "P": P.copy()}, saveqr=saveqr)
    c_ref, kappa, rel_rms, R = LO.lsqfit_qr(ref, db.configs, w, E0=None, P=P)
    assert np.linalg.norm(info["c"] - c_ref) <= C_RTOL * np.linalg.norm(c_ref)
    assert abs(info["rel_rms"] - rel_rms) <= 1e-8 * rel_rms
    assert abs(info["cond_R"] - kappa) <= 1e-6 * kappa
    assert np.abs(np.abs(saveqr["R"]) - np.abs(R)).max() <= 1e-8 * np.abs(R).max()       # factors agree up to row signs
    assert "P" not in info                                                                # `delete!(solver, "P")`
    with pytest.raises(NotImplementedError):
        ipf.lsqfit(db, weights=dict(w), E0=0.0, solver={"solver": "svd", "svd_ndiscard": 0, "P": P.copy()})
