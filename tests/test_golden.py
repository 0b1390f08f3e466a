"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
the oracle must keep reproducing them (CPU), and the CUDA path must match them (GPU)."""
import os

import numpy as np
import pytest

import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200.lsq_db import LsqDB, _alloc_rows
import ipf_oracle as O
import lsq_oracle as LO
from util import PSI_RTOL, psi_errors, rows_for

import importlib.util
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


def _load(name):
    return np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))


@pytest.mark.parametrize("name", G.CASES)
def test_oracle_reproduces_golden(name):
    g = _load(name)
    B, configs = G.case_basis(name), G.case_configs(name)
    n = _alloc_rows(configs)
    r = rows_for(configs)
    Psi = O.assemble(B, configs, r["E"], r["F"], r["V"], n)
    assert Psi.shape == g["Psi"].shape
    assert psi_errors(Psi, g["Psi"], configs) <= 1e-12        # libm differences only
    I, J, S, R = O.neighbours(g["X0"], g["cell0"], (True, True, True), B.rcut)
    assert np.array_equal(I, g["nl_I"]) and np.array_equal(J, g["nl_J"]) and np.array_equal(S, g["nl_S"])
    assert np.array_equal(R, g["nl_R"])
    c, kappa, rel, _ = LO.lsqfit_qr(Psi, configs, G.WEIGHTS, E0=-4.0)
    assert np.linalg.norm(c - g["c"]) <= 1e-7 * np.linalg.norm(g["c"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", G.CASES)
def test_cuda_matches_golden(ctx, name):
    from ipfitting_jl_b200 import _capi
    g = _load(name)
    B, configs = G.case_basis(name), G.case_configs(name)
    db = LsqDB("", B, configs, ctx=ctx)
    assert psi_errors(db.Psi, g["Psi"], configs) <= PSI_RTOL
    I, J, S, R = _capi.neighbours(ctx, g["X0"], g["cell0"], (True, True, True), B.rcut)
    assert np.array_equal(I, g["nl_I"]) and np.array_equal(J, g["nl_J"]) and np.array_equal(S, g["nl_S"])
    assert np.array_equal(R.view(np.uint64), g["nl_R"].view(np.uint64))
    IP, info = ipf.lsqfit(db, weights=dict(G.WEIGHTS), E0=-4.0, Vref=ipf.OneBody(-4.0))
    kappa = float(g["kappa"])
    assert np.linalg.norm(info["c"] - g["c"]) <= max(1e-8, 1e2 * kappa * 2.2e-16) * np.linalg.norm(g["c"])
    assert abs(info["rel_rms"] - float(g["rel_rms"])) <= 1e-8 * float(g["rel_rms"])
